"""Text batches through the C ABI: records parsed on the device (bgr_slot_upload_text, replacing getReads,
aligner.cpp:46-117, for text of the regular shape) and the three output streams formatted there
(bgr_slot_format_text, replacing the output half of alignPartGreedy + printPath, alignerGreedy.cpp:393-428,
aligner.cpp:600-609). Expected streams are written here from the oracle's per-read results and getReads' filters."""
import random

import pytest

import synth_small as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bg():
    import bgreat_b200
    return bgreat_b200


def _kept(read, k, fastq):
    return len(read) > 2 and (fastq or len(read) > k) and set(read) <= set("ACGTN")


def _expected(o, k, headers, reads, mode, fastq, split):
    out = ["", "", ""]
    n = 0
    for h, r in zip(headers, reads):
        if not _kept(r, k, fastq):
            continue
        n += 1
        of, rc, path = o.align(r, mode)
        if path:
            out[0] += h + "\n" + "".join("%d." % x for x in path) + "\n"
        elif split and not of:
            out[2] += h + "\n" + r + "\n"
        else:
            out[1] += h + "\n" + r + "\n"
    return n, [x.encode() for x in out]


def _case(seed, k=31, n=3000):
    rng = random.Random(seed)
    unitigs, haps = S.make_graph(rng, k, genome_len=30000, mean_unitig=70, n_variants=200)
    reads = S.make_reads(rng, haps, n, 150, sub_rate=0.01, n_rate=0.0005)
    # what getReads drops: too short, foreign characters (lower case, IUPAC), empty
    reads += ["ACGTACGTAC", "AC", "", "acgt" * 20, "ACGTRYACGT" * 8, "ACGT" * 20 + "\r"]
    reads += S.make_reads(rng, haps, 200, 40, sub_rate=0.02) + S.make_reads(rng, haps, 50, 400, sub_rate=0.01)
    rng.shuffle(reads)
    return unitigs, reads


@pytest.mark.parametrize("fastq", [False, True], ids=["fasta", "fastq"])
@pytest.mark.parametrize("mode,split", [(O.GREEDY, False), (O.GREEDY, True), (O.EXHAUSTIVE, False)])
def test_text_batch_equals_oracle_streams(bg, fastq, mode, split):
    k = 31
    unitigs, reads = _case(3 + int(fastq))
    if not fastq:
        reads = [r for r in reads if not r.startswith(">")]
    headers = [("@" if fastq else ">") + "read_%d some text" % i for i in range(len(reads))]
    if fastq:
        text = "".join("%s\n%s\n+\n%s\n" % (h, r, "I" * len(r)) for h, r in zip(headers, reads))
    else:
        text = "".join("%s\n%s\n" % (h, r) for h, r in zip(headers, reads))
    al = bg.Aligner(k, 2, 2)
    al.index_unitigs(unitigs)
    o = O.Oracle(k, 2, 2)
    o.index_unitigs(unitigs)
    want_n, want = _expected(o, k, headers, reads, mode, fastq, split)
    regular, n, got = al.map_text(text, fastq, mode, split)
    assert regular and n == want_n
    assert got == want
    c, oc = al.counters(), o.counters()
    for name in ("reads", "aligned", "not_aligned", "no_overlap"):
        assert c[name] == oc[name], name
    al.close()


def test_text_batch_correction_mode(bg):
    k = 31
    unitigs, reads = _case(11, n=1500)
    reads = [r for r in reads if _kept(r, k, False)]
    headers = [">r%d" % i for i in range(len(reads))]
    text = "".join("%s\n%s\n" % (h, r) for h, r in zip(headers, reads))
    al = bg.Aligner(k, 2, 2)
    al.index_unitigs(unitigs)
    o = O.Oracle(k, 2, 2)
    o.index_unitigs(unitigs)
    import bgreat_b200._abi as A
    regular, n, got = al.map_text(text, False, O.GREEDY, False, flags=A.P_CORRECT)
    assert regular and n == len(reads)
    want = ["", ""]
    for h, r in zip(headers, reads):
        cr = o.correct(r)
        if cr is not None:
            want[0] += h + "\n" + (cr.decode() if isinstance(cr, bytes) else cr) + "\n"
        else:
            want[1] += h + "\n" + r + "\n"
    assert got[0] == want[0].encode() and got[1] == want[1].encode()
    al.close()


@pytest.mark.parametrize("text,fastq", [
    (">a\nACGT\nACGT\n>b\nACGT\n", False),        # a continuation line
    (">a\nACGTACGT", False),                      # no final newline
    ("a\nACGTACGT\n", False),                     # header without '>'
    (">a\n>b\nACGT\n>c\nACGT\n", False),          # header where a sequence should be
    (">a\nACGT\n>b\n", False),                    # header without a sequence line
    ("@a\nACGT\n+\nIIII\n@b\nACGT\n+\n", True),   # partial record
])
def test_irregular_text_is_reported_not_guessed(bg, text, fastq):
    al = bg.Aligner(4, 2, 2)
    al.index_unitigs(["ACGTACGTTGCA", "TTGCAACGG"])
    regular, n, got = al.map_text(text, fastq)
    assert not regular and n == 0
    al.close()


def test_empty_and_all_dropped_text(bg):
    al = bg.Aligner(31, 2, 2)
    al.index_unitigs(["ACGT" * 20])
    assert al.map_text("", False) == (True, 0, [b"", b"", b""])
    assert al.map_text(">a\nACGT\n>b\nNNNNNNNN\n", False) == (True, 0, [b"", b"", b""])  # shorter than k
    al.close()
