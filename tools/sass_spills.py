"""Where a kernel spills: STL / LDL instructions of one function of an object file, grouped by source line (-lineinfo).

    python tools/sass_spills.py /tmp/schur.o k_eliminateILi1ELi0ELb0
"""
import collections, re, subprocess, sys, tempfile, os

obj, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
inside, cur = False, None
cnt, ninstr = collections.Counter(), 0
for line in sass.splitlines():
    if line.startswith("//---") and ".text." in line:
        inside = pat in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.search(r"^\s+/\*[0-9a-f]{4,}\*/", line):
        ninstr += 1
        if re.search(r"\b(STL|LDL)\b", line):
            cnt[(cur, "STL" if "STL" in line else "LDL")] += 1
print("instructions:", ninstr)
for k, v in sorted(cnt.items(), key=lambda x: (x[0][0][0], x[0][0][1])):
    print(k[0][0], k[0][1], k[1], v)
