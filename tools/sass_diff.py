#!/usr/bin/env python
"""Per-kernel SASS comparison of two builds of the library (instruction text, addresses stripped, anonymous-
namespace hashes normalised): which kernels of `new` differ from, or are missing in, `old`.  Used to show that a
refactor or an addition left the kernels already measured on the GPU byte-identical.

    git worktree add build/old_tree <commit> && make -C build/old_tree -j8
    python tools/sass_diff.py build/old_tree/build build

Caveat: ptxas is not deterministic for every kernel -- compiling the unchanged dwpw_tc.cu six times in a row gave two
different uniform-register allocations (UR8 vs UR10 in the same instruction) -- so a CHANGED line for a file whose
source did not change has to be confirmed by rebuilding that object a few times.
"""
import glob
import hashlib
import os
import re
import subprocess
import sys


def norm(name):
    return re.sub(r'_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}', 'ANON', name)


def kernels(obj):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True).stdout
    res, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            if name:
                res[name] = hashlib.md5('\n'.join(body).encode()).hexdigest()
            name, body = norm(m.group(1)), []
        elif name:
            text = re.sub(r'/\*[0-9a-fx]+\*/', '', line).strip()
            if text:
                body.append(text)
    if name:
        res[name] = hashlib.md5('\n'.join(body).encode()).hexdigest()
    return res


def main(old_dir, new_dir):
    same = changed = added = 0
    for old in sorted(glob.glob(os.path.join(old_dir, '*.o'))):
        new = os.path.join(new_dir, os.path.basename(old))
        if not os.path.exists(new):
            print('MISSING FILE', os.path.basename(old))
            continue
        fo, fn = kernels(old), kernels(new)
        for k, v in fo.items():
            if fn.get(k) != v:
                print('CHANGED' if k in fn else 'MISSING', os.path.basename(old), k[:100])
                changed += 1
            else:
                same += 1
        added += len(set(fn) - set(fo))
    new_files = sorted(set(map(os.path.basename, glob.glob(os.path.join(new_dir, '*.o')))) -
                       set(map(os.path.basename, glob.glob(os.path.join(old_dir, '*.o')))))
    print('identical kernels: %d, changed or missing: %d, added to existing files: %d, new object files: %s'
          % (same, changed, added, new_files))
    return 1 if changed else 0


if __name__ == '__main__':
    sys.exit(main(sys.argv[1], sys.argv[2]))
