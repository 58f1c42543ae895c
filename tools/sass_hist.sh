# opcode histogram of one kernel in an object file: tools/sass_hist.sh <obj> <mangled-name-substring>
cuobjdump -sass "$1" | awk -v pat="$2" '/Function : /{f = index($0, pat) > 0} f' > /tmp/k.sass
echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/k.sass)"
grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/k.sass | awk '{print $2}' | sed 's/;//' | sort | uniq -c | sort -rn | head -${3:-16}
