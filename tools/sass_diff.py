#!/usr/bin/env python
"""Which kernels changed?  Builds the CUDA sources of a git revision (default HEAD) next to the working tree's
libteal_b200.so and compares the SASS of every kernel, instruction by instruction (nvcc + cuobjdump, no GPU needed).

    python tools/sass_diff.py [REV]

"identical": same instruction stream; "imm-only": only immediates / offsets differ (e.g. a struct grew); "DIFF": the
code changed -- a kernel that has been validated on hardware must not show up as DIFF without being re-validated."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    d, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "_ANON_", m.group(1))
            d[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            d[cur].append(re.sub(r"\s+", " ", m.group(1)))
    return d


def demangle(name):
    r = subprocess.run(["c++filt", name.replace("_ANON_", "")], capture_output=True, text=True).stdout.strip()
    return (r or name)[:110]


def main():
    rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
    sys.path.insert(0, ROOT)
    from teal_b200 import build
    new_lib = build.build()
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "teal_b200/csrc", "teal_b200/build.py", "include"],
                             capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        srcs = sorted(os.path.join(tmp, "teal_b200", "csrc", f) for f in os.listdir(os.path.join(tmp, "teal_b200", "csrc")) if f.endswith(".cu"))
        old_lib = os.path.join(tmp, "old.so")
        subprocess.run(["nvcc", *build.NVCC_FLAGS, "-shared", "-o", old_lib, *srcs], check=True)
        old, new = kernels(old_lib), kernels(new_lib)
    counts = collections.Counter()
    renamed = set()
    for k, a in old.items():
        if k not in new:
            twin = next((n for n, b in new.items() if n not in old and n not in renamed and b == a), None)
            if twin is not None:  # e.g. a template gained a defaulted parameter: same code under a new mangled name
                renamed.add(twin)
                counts["identical (renamed)"] += 1
                continue
            status = "REMOVED"
        elif a == new[k]:
            status = "identical"
        elif [re.sub(r"0x[0-9a-f]+", "", x) for x in a] == [re.sub(r"0x[0-9a-f]+", "", x) for x in new[k]]:
            status = "imm-only"
        else:
            status = "DIFF"
        counts[status] += 1
        if status != "identical":
            print(f"{status:9s} {demangle(k)}  ({len(a)} -> {len(new.get(k, []))} instructions)")
    for k in new:
        if k not in old and k not in renamed:
            counts["new"] += 1
            print(f"{'new':9s} {demangle(k)}  ({len(new[k])} instructions)")
    print(dict(counts))


if __name__ == "__main__":
    main()
