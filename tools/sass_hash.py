#!/usr/bin/env python
"""MD5 of the SASS of every kernel in object files / libraries, to prove that a refactoring or a comment-only change
left the machine code of a measured kernel untouched (no GPU needed):
    python tools/sass_hash.py a.o b.o"""
import hashlib
import re
import subprocess
import sys


def hashes(obj):
    s = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    d = {}
    for q in re.split(r"\n\t\tFunction : ", s)[1:]:
        name = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_", "", q.split("\n")[0])
        body = "\n".join(q.split("\n\t\t.......")[0].split("\n")[1:])
        d[name] = hashlib.md5(body.encode()).hexdigest()
    return d


if __name__ == "__main__":
    tabs = [hashes(o) for o in sys.argv[1:]]
    for name in sorted(set().union(*tabs)):
        hs = [t.get(name, "-") for t in tabs]
        print(("same " if len(set(hs)) == 1 else "DIFF ") + name[:70] + "  " + " ".join(h[:10] for h in hs))
