"""CPU-side parity gates (no GPU needed):
 * the oracle (the unmodified reference built under oracle/_ref) reproduces the committed golden vectors;
 * the per-read / per-position device functions, single-stepped on the CPU by tests/tools/rv_dump.cpp
   (`--backend sim`, a test-only harness), reproduce them too.
"""
import os

import pytest

import cases
import dumpcmp
from conftest import run, unpack_golden


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_reproduces_golden(built, ref_tools, name, tmp_path):
    if ref_tools is None:
        pytest.skip("oracle/_ref not built here")
    c = cases.CASES[name]
    d = cases.generate(name)
    got = str(tmp_path / "ref.txt")
    env = dict(os.environ, RV_DUMP=got, RV_DUMP_STAGES=c["stages"])
    run([ref_tools["ref_dump"]] + c["ref_args"](d), env=env)
    want = unpack_golden(name, "dump.txt", str(tmp_path / "golden.txt"))
    n, problems = dumpcmp.compare(want, got, None, rel_tol=0.0)
    assert not problems, problems[:5]


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_device_logic_single_stepped_matches_golden(built, name, tmp_path):
    c = cases.CASES[name]
    cases.generate(name)
    got = str(tmp_path / "sim.txt")
    run(cases.dump_cmd(name, "sim", got, c["stages"]))
    want = unpack_golden(name, "dump.txt", str(tmp_path / "golden.txt"))
    # integer / string fields bit-exact, floating-point fields within 1e-9 relative (north_star)
    n, problems = dumpcmp.compare(want, got, c["exact_stages"])
    assert n > 0 or name == "edge_empty"
    assert not problems, problems[:5]


def test_smoke_golden_matches_single_stepped_device_logic(built, tmp_path):
    """tests/golden/smoke_C.txt.gz is what __graft_entry__.smoke() falls back to on a box without oracle/_ref."""
    import gzip
    import __graft_entry__ as g
    work = str(tmp_path / "smoke")
    os.makedirs(work)
    g._smoke_dataset(work)
    got = str(tmp_path / "sim.txt")
    run([os.path.join(cases.BUILD, "rv_dump"), "--backend", "sim", "--fasta", os.path.join(work, "ref.fa"), "--bam",
         os.path.join(work, "S.bam"), "--chr", "chrS1", "--region", "1301-13300", "--out", got, "--stages", "C"])
    want = str(tmp_path / "golden.txt")
    with gzip.open(os.path.join(cases.ROOT, "tests", "golden", "smoke_C.txt.gz"), "rt") as f, open(want, "w") as o:
        o.write(f.read())
    n, problems = dumpcmp.compare(want, got, ["C."])
    assert n > 0 and not problems, problems[:5]
