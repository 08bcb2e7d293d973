"""Field-wise comparison of two table dumps (oracle/ref_dump.cpp vs tests/tools/rv_dump.cpp).

Integer and string fields must be identical; floating-point fields must agree within REL_TOL
relative (north_star: 1e-9; the oracle itself is compiled with -ffast-math, CMakeLists.txt:27, so
its doubles are not always the correctly-rounded quotient).
"""
import math
import sys

REL_TOL = 1e-9


def _is_float(tok):
    return any(c in tok for c in ".eE") and tok.replace(".", "").replace("-", "").replace("e", "").replace("E", "").replace("+", "").isdigit()


def fields_equal(a, b, rel_tol=REL_TOL):
    if a == b:
        return True
    try:
        fa, fb = float(a), float(b)
    except ValueError:
        return False
    if not (_is_float(a) or _is_float(b)):
        return False
    if math.isnan(fa) or math.isnan(fb):
        return math.isnan(fa) and math.isnan(fb)
    return abs(fa - fb) <= rel_tol * max(abs(fa), abs(fb))


def load(path, prefixes=None):
    rows = {}
    with open(path) as f:
        region = ""
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            t = line.split("\t")
            if t[0] == "REGION":
                region = "\t".join(t[1:])
                continue
            if prefixes and not any(t[0].startswith(p) for p in prefixes):
                continue
            # key: region + tag + position-ish identifying fields (everything non-numeric up to the counters)
            tag = t[0]
            if tag.endswith(".COV") or tag.endswith("MAXRL"):
                nkey = 2 if tag.endswith(".COV") else 1
            elif tag.endswith(".SCNT") or tag.endswith(".SCSEQ"):
                nkey = 5
            elif tag.endswith(".SC5") or tag.endswith(".SC3"):
                nkey = 2
            else:
                nkey = 3
            key = (region, tag) + tuple(t[1:nkey])
            rows.setdefault(key, []).append(t[nkey:])
    return rows


def compare(ref_path, got_path, prefixes=None, rel_tol=REL_TOL, max_report=20):
    ref, got = load(ref_path, prefixes), load(got_path, prefixes)
    problems = []
    n = 0
    for k in sorted(set(ref) | set(got)):
        n += 1
        if k not in got:
            problems.append(("missing", k, ref[k]))
            continue
        if k not in ref:
            problems.append(("extra", k, got[k]))
            continue
        ra, ga = ref[k], got[k]
        if len(ra) != len(ga):
            problems.append(("count", k, (ra, ga)))
            continue
        for r, g in zip(ra, ga):
            if len(r) != len(g) or not all(fields_equal(x, y, rel_tol) for x, y in zip(r, g)):
                problems.append(("differs", k, (r, g)))
    return n, problems


if __name__ == "__main__":
    prefixes = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    n, problems = compare(sys.argv[1], sys.argv[2], prefixes)
    kinds = {}
    for p in problems:
        kinds[p[0]] = kinds.get(p[0], 0) + 1
    print(f"{n} keys compared, {len(problems)} problems {kinds}")
    for p in problems[:int(sys.argv[4]) if len(sys.argv) > 4 else 20]:
        print(p)
    sys.exit(1 if problems else 0)
