#!/bin/bash
# developer tool: static SASS instruction count per kernel in an object file
cuobjdump -sass "$1" | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/{cnt[name]++} END{for(n in cnt) print cnt[n], substr(n,25,60)}' | sort -k2
