#!/bin/bash
# usage: scripts/sass_hist.sh lib.so mangled_function_substring
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{f=index($0,pat)>0} f' > /tmp/fn.sass
echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]+\*/\s+[A-Z@!]' /tmp/fn.sass)"
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)" /tmp/fn.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-16}
