#!/usr/bin/env python
"""tools/sass_diff.py OLD.o NEW.o [--map OLDSUBSTR=NEWSUBSTR] -- are the kernels of two objects the same machine code?

Compares `cuobjdump -sass` function by function, ignoring encodings and -lineinfo comments.  Used to prove that adding
an A/B variant (a new template parameter, an env-gated branch) leaves every kernel that was validated on the GPU
bit-identical; `--map` renames mangled names when a defaulted template parameter was appended."""
import re
import subprocess
import sys


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        line = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip()
        if cur and line and not line.startswith("//") and not line.startswith("."):
            d[cur].append(re.sub(r"\s+", " ", line))
    return d


def main():
    old, new = functions(sys.argv[1]), functions(sys.argv[2])
    maps = [a.split("=", 1) for a in sys.argv[3:] if "=" in a and not a.startswith("--")]
    same, diff, missing = 0, [], []
    for k, v in old.items():
        k2 = k
        if k2 not in new:                         # renamed (a defaulted template parameter was appended)?
            for a, b in maps:
                k2 = k2.replace(a, b)
        if k2 not in new:
            missing.append(k)
        elif new[k2] == v:
            same += 1
        else:
            diff.append(k2)
    print(f"{same} identical, {len(diff)} different, {len(missing)} missing, {len(new) - same - len(diff)} new-only")
    for k in diff:
        print("DIFF", k)
    for k in missing:
        print("MISSING", k)
    sys.exit(1 if diff or missing else 0)


if __name__ == "__main__":
    main()
