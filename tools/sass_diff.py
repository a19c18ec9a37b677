"""Compare two `cuobjdump -sass` dumps function by function (used to prove that adding an opt-in kernel or a
template parameter left the SASS of the default-path kernels untouched).  usage: sass_diff.py old.sass new.sass"""
import re
import sys


def functions(path):
    out, name, body = {}, None, []
    for line in open(path):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name and "/*" in line:
            # drop the address / encoding columns, keep the instruction text
            t = re.sub(r"/\*\s*(0x)?[0-9a-f]{4,}\s*\*/", "", line).strip()
            # c[0x4][..] holds relocated addresses (printf format strings): the offset moves when any string is added
            t = re.sub(r"c\[0x4\]\[0x[0-9a-f]+\]", "c[0x4][reloc]", t)
            if t:
                body.append(t)
    if name:
        out[name] = body
    return out


def main():
    a, b = functions(sys.argv[1]), functions(sys.argv[2])
    changed = [k for k in a if k in b and a[k] != b[k]]
    print("functions: %d old, %d new; removed %d, added %d, changed %d" % (
        len(a), len(b), len(set(a) - set(b)), len(set(b) - set(a)), len(changed)))
    for k in sorted(set(b) - set(a)):
        print("  added  ", k)
    for k in sorted(set(a) - set(b)):
        print("  removed", k)
    for k in changed:
        print("  changed", k)
    return 1 if changed or set(a) - set(b) else 0


if __name__ == "__main__":
    sys.exit(main())
