"""Static code size per source line of one kernel: `nvdisasm -g` of the extracted cubin, instructions counted under the
last '//## File … line N' marker.  usage: sass_lines.py <nvdisasm.txt> <function-substring> [top]"""
import re, sys, collections
txt = open(sys.argv[1]).read().split("\n")
fn = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
inside = False
cur = None
cnt = collections.Counter()
total = 0
for ln in txt:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        inside = fn in m.group(1)
        continue
    if not inside:
        continue
    if re.match(r"\s*\$_Z\S+:", ln):          # a device function placed inside the kernel's section
        inside = len(sys.argv) > 4 and sys.argv[4] in ln
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        cnt[cur] += 1
        total += 1
print("instructions", total, "bytes", total * 16)
src = {}
for (f, l), c in cnt.most_common(top):
    if f not in src:
        try:
            src[f] = open("/root/repo/beluga.jl_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print("%5d  %-18s:%-5d %s" % (c, f, l, text))
