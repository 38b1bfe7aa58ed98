#!/bin/bash
# usage: tools/sassmix.sh <object-or-so> <function-name-substring>
# Prints the SASS opcode histogram of the first function whose mangled name contains the substring.
obj=$1; pat=$2
fn=$(cuobjdump -sass "$obj" | grep "Function :" | grep -- "$pat" | head -1 | awk '{print $3}')
[ -z "$fn" ] && { echo "no function matching $pat"; exit 1; }
cuobjdump -sass -fun "$fn" "$obj" > /tmp/sassmix.sass
echo "# $fn: $(grep -cE '^\s+/\*[0-9a-f]{4,}\*/' /tmp/sassmix.sass) instructions"
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9] )?[A-Z0-9_.]+" /tmp/sassmix.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-25}
