"""interpolation shim (test / baseline infrastructure); see oracle/shims/README.md."""
