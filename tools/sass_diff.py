"""Per-kernel SASS comparison of two builds of libplvs_b200.so (cuobjdump -sass, addresses stripped): which kernels changed between two commits.
Usage: python tools/sass_diff.py OLD.so NEW.so"""
import subprocess, re, sys, hashlib
def funcs(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res = {}
    cur = None
    for line in out.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+_cu_[0-9a-f]+", "anon", name)
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = name.split("(")[0]; res[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            # strip the address column and the encoding comment
            ins = re.sub(r"/\*[0-9a-f]+\*/", "", line).strip()
            res[cur].append(ins)
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() + ":%d" % len(v) for k, v in res.items()}
a = funcs(sys.argv[1]); b = funcs(sys.argv[2])
same = [k for k in a if k in b and a[k] == b[k]]
diff = [k for k in a if k in b and a[k] != b[k]]
print("same", len(same), "different", len(diff), "only old", [k for k in a if k not in b], "only new", [k for k in b if k not in a])
for k in diff: print("DIFF", k, a[k], b[k])
