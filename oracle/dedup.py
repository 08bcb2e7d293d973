"""TEST INFRASTRUCTURE ONLY (oracle): literal restatement of the reference's -t duplicate filter and of the
order-free formulation the CUDA path uses, so that tests can check that the two agree on arbitrary record streams.

Reference: RecordPreprocessor::next_record, /root/reference/src/recordPreprocessor.cpp:153-176 (the running set
`duplicates`, cleared whenever a record's start differs from `firstMatchingPosition`), getMateReferenceName :80-90.
Parity pinned by the goldens `dedup_t` / `dedup_t_F500` made from the reference itself (tests/golden/make_golden.py);
this module only pins the equivalence of the two formulations.

A record is a dict: pos (1-based start), mpos (1-based mate start), flag, tid, mtid, cigar (tuple of u32 ops).
Records are the ones that already passed the filters running before the -t block (:121-146), in BAM order.
"""


def mate_reference_name(r):
    """recordPreprocessor.cpp:80-90"""
    if not (r["flag"] & 0x1):
        return "*"
    if r["tid"] == r["mtid"]:
        return "="
    return "chr%d" % r["mtid"]


def cigar_string(r):
    return "".join("%d%s" % (c >> 4, "MIDNSHP=XB"[c & 15]) for c in r["cigar"])


def running_set_filter(records):
    """The reference's algorithm, statement by statement.  Returns the kept flags."""
    duplicates = set()
    first_matching_position = 0
    kept = []
    for r in records:
        if r["pos"] != first_matching_position:
            duplicates.clear()
        if r["mpos"] < 10:
            key = "%d-%s-%d" % (r["pos"], mate_reference_name(r), r["mpos"])
            if key in duplicates:
                kept.append(False)
                continue
            duplicates.add(key)
            first_matching_position = r["pos"]
        elif (r["flag"] & 0x1) and (r["flag"] & 0x4):
            key = "%d-%s" % (r["pos"], cigar_string(r))
            if key in duplicates:
                kept.append(False)
                continue
            duplicates.add(key)
            first_matching_position = r["pos"]
        kept.append(True)
    return kept


def key_kind(r):
    if r["mpos"] < 10:
        return 1
    if (r["flag"] & 0x1) and (r["flag"] & 0x4):
        return 2
    return 0


def scan_back_filter(records):
    """What rvk::is_duplicate_read (rabbitvar_b200/csrc/kernels/rv_core.cuh) computes, one record at a time and
    independent of the others' results: a record is dropped iff an earlier record with the same start has the same key."""
    kept = []
    for i, me in enumerate(records):
        kind = key_kind(me)
        dup = False
        if kind:
            j = i - 1
            while j >= 0 and records[j]["pos"] == me["pos"]:
                o = records[j]
                j -= 1
                if key_kind(o) != kind:
                    continue
                if kind == 1:
                    if o["mpos"] != me["mpos"]:
                        continue
                    mp, op = bool(me["flag"] & 1), bool(o["flag"] & 1)
                    if mp != op:
                        continue
                    if mp:
                        ms, os_ = me["tid"] == me["mtid"], o["tid"] == o["mtid"]
                        if ms != os_ or (not ms and me["mtid"] != o["mtid"]):
                            continue
                    dup = True
                    break
                if o["cigar"] == me["cigar"]:
                    dup = True
                    break
        kept.append(not dup)
    return kept
