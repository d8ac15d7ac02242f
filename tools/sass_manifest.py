#!/usr/bin/env python
"""tools/sass_manifest.py --write FILE | --check FILE -- fingerprint of every kernel's machine code in
ac_dsp_b200/lib/obj/*.o (cuobjdump -sass, encodings and -lineinfo comments ignored, sha256 per kernel).

Written after a GPU run that passed (`--write`), checked after host-only edits (`--check`): proves the device code in
the tree is instruction for instruction the one that was validated on the B200, and lists what changed otherwise."""
import glob
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_diff import functions  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def manifest():
    out = {}
    for obj in sorted(glob.glob(os.path.join(ROOT, "ac_dsp_b200", "lib", "obj", "*.cu.o"))):
        for name, lines in functions(obj).items():
            out[f"{os.path.basename(obj)[:-5]}::{name}"] = hashlib.sha256("\n".join(lines).encode()).hexdigest()[:20]
    return out


def main():
    mode, path = sys.argv[1], sys.argv[2]
    cur = manifest()
    if mode == "--write":
        json.dump(cur, open(path, "w"), indent=0, sort_keys=True)
        print(f"{len(cur)} kernels -> {path}")
        return
    old = json.load(open(path))
    changed = [k for k in old if k in cur and cur[k] != old[k]]
    missing = [k for k in old if k not in cur]
    new = [k for k in cur if k not in old]
    print(f"{len(old) - len(changed) - len(missing)} identical, {len(changed)} changed, {len(missing)} missing, {len(new)} new")
    for tag, lst in (("CHANGED", changed), ("MISSING", missing), ("NEW", new)):
        for k in lst:
            print(tag, k)
    sys.exit(1 if changed or missing else 0)


if __name__ == "__main__":
    main()
