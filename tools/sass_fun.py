#!/usr/bin/env python
"""Print the SASS of the first kernel whose mangled name contains the given substring (no encodings).
usage: python tools/sass_fun.py <lib.so|cubin> <substring>"""
import re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for b in re.split(r"\n\s*Function : ", txt)[1:]:
    name = b.split("\n", 1)[0].strip()
    if sys.argv[2] in name:
        print("//", name)
        for line in b.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
            if m:
                print(m.group(1), m.group(2).strip())
        break
