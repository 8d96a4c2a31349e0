#!/usr/bin/env python
"""Generate the committed golden vectors from the COMPILED, UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref built by oracle/Makefile):
    python tests/golden/make_golden.py
Writes tests/golden/{harness_cases,cli_cases,parse_cases}.json. The GPU box has no
/root/reference: tests there only read these JSON files.

  harness_cases: per-read results of Aligner::alignReadGreedy / alignReadExhaustive /
                 alignReadGreedyPath / alignReadExhaustivePath (oracle/ref_harness.cpp)
  cli_cases    : files + stdout of the reference `bgreat` binary
  parse_cases  : (header, read) pairs Aligner::getReads yields
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
import synth_small as S  # noqa: E402

REF_DIR = "/root/reference"
MODES = [O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH, O.EXHAUSTIVE_PATH]


def read_fasta_seqs(path):
    return [l.strip() for l in open(path) if l.strip() and not l.startswith(">")]


def harness_case(name, k, m, effort, unitigs, reads):
    with tempfile.TemporaryDirectory() as d:
        uf = os.path.join(d, "u.fa")
        S.write_unitigs(uf, unitigs)
        results = {}
        for mode in MODES:
            res, counters = O.ref_harness_align(uf, k, m, effort, mode, reads, d)
            results[O.MODE_NAMES[mode]] = dict(
                reads=[[int(of), int(rc)] + p for (of, rc, p) in res], counters=counters)
    return dict(name=name, k=k, m=m, effort=effort, unitigs=unitigs, reads=reads, results=results)


def known_answer_cases():
    cases = []
    # K1/K2: the bundled toy data (reference reads.fa / unitig.fa), k=4
    u = read_fasta_seqs(os.path.join(REF_DIR, "unitig.fa"))
    r = read_fasta_seqs(os.path.join(REF_DIR, "reads.fa"))
    for m in (0, 1, 2):
        cases.append(harness_case("K1_bundled_m%d" % m, 4, m, 2, u, r))
    # K3/K9: double-counted right-overlap mismatches; effort changes the result
    k3u = ["AAATC", "ATCGGA", "GGATT"]
    for m in (0, 1, 2, 3):
        for e in (1, 2, 3):
            cases.append(harness_case("K3_m%d_e%d" % (m, e), 4, m, e, k3u, ["AAATCGCATT", "AAATCGGATT", "AATCCGATTT"]))
    # K6: five unitigs beginning ATC -> slot 4 overwritten
    k6u = ["ATCAA", "ATCCC", "ATCGG", "ATCTT", "ATCAG", "GGATC", "TTGATC"]
    k6r = ["GGATCAA", "GGATCCC", "GGATCGG", "GGATCTT", "GGATCAG", "TTGATCTT", "AAGATCC", "CTGATCC"]
    for m in (0, 1, 2):
        cases.append(harness_case("K6_slot_overflow_m%d" % m, 4, m, 2, k6u, k6r))
    # K8: N characters
    for m in (0, 1, 2):
        cases.append(harness_case("K8_N_m%d" % m, 4, m, 2, k3u,
                                  ["ANATCGGATT", "AAATCGGNTT", "AAATCGNATT", "NAATCGGATT", "AAATCGGATN", "NNNNNNNNNN",
                                   "AAANCGGATT"]))
    # K10: hairpin unitig (begins AAC, ends rc(AAC)=GTT)
    cases.append(harness_case("K10_hairpin", 4, 0, 2, ["AACTGTT", "CCAAC", "AACGG"], ["CCAACTGTT", "CCAACGG", "AACAGTTGG"]))
    # palindromic border (k-1 = 4: ACGT), both orientations around it
    cases.append(harness_case("palindrome_k5", 5, 1, 2, ["GGACGT", "ACGTCC", "TTACGT", "ACGTAA"],
                              ["GGACGTCC", "TTACGTAA", "GGACGTAA", "TTACGTCC", "GGACGTCA"]))
    # reads of length k-1, k, k+1 straight through the per-read API
    cases.append(harness_case("short_reads_k4", 4, 1, 2, k3u, ["AAT", "AAAT", "AAATC", "ATCG", "TCGG", "GGA", "ATCGG"]))
    return cases


def cli_case(name, args, files, k_outputs=("paths", "notAligned.fa")):
    with tempfile.TemporaryDirectory() as d:
        for fn, content in files.items():
            with open(os.path.join(d, fn), "w", newline="") as f:
                f.write(content)
        stdout = O.ref_bgreat(args, d)
        outs = {}
        for fn in k_outputs:
            p = os.path.join(d, fn)
            outs[fn] = open(p, newline="").read() if os.path.exists(p) else None
        created = sorted(set(os.listdir(d)) - set(files))
    # timing lines depend on the wall clock
    keep = [l for l in stdout.splitlines()
            if not l.startswith(("Indexing in seconds", "Reads/seconds", "Mapping in seconds"))]
    return dict(name=name, args=args, files=files, outputs=outs, created=created, stdout=keep)


def cli_cases():
    cases = []
    unitig_fa = open(os.path.join(REF_DIR, "unitig.fa")).read()
    reads_fa = open(os.path.join(REF_DIR, "reads.fa")).read()
    base = {"unitig.fa": unitig_fa, "reads.fa": reads_fa}
    cases.append(cli_case("K1", ["-r", "reads.fa", "-k", "4", "-g", "unitig.fa", "-m", "2", "-t", "1"], base))
    cases.append(cli_case("K1_default_unitig_name", ["-r", "reads.fa", "-k", "4", "-m", "2"], base))
    cases.append(cli_case("K2_brute", ["-r", "reads.fa", "-k", "4", "-g", "unitig.fa", "-m", "2", "-b"], base))
    cases.append(cli_case("K1_f_a", ["-r", "reads.fa", "-k", "4", "-g", "unitig.fa", "-f", "out.paths", "-a", "out.na"],
                          base, k_outputs=("out.paths", "out.na", "paths", "notAligned.fa")))
    cases.append(cli_case("K1_o_p_inert", ["-r", "reads.fa", "-k", "4", "-g", "unitig.fa", "-o", "nov.fa", "-p"], base,
                          k_outputs=("paths", "notAligned.fa", "nov.fa")))
    cases.append(cli_case("K1_two_files", ["-r", "reads.fa,reads.fa", "-k", "4", "-g", "unitig.fa"], base))
    cases.append(cli_case("no_args", [], base))
    # K4: FASTQ with / without trailing newline (phantom read)
    k3u = ">1\nAAATC\n>2\nATCGGA\n>3\nGGATT\n"
    fq = "@a\nAAATCGGATT\n+\nIIIIIIIIII\n@b\nAAATCGCATT\n+\nIIIIIIIIII"
    cases.append(cli_case("K4_fastq_trailing_nl", ["-r", "r.fq", "-q", "-k", "4", "-g", "u.fa", "-m", "1"],
                          {"u.fa": k3u, "r.fq": fq + "\n"}))
    cases.append(cli_case("K4_fastq_no_trailing_nl", ["-r", "r.fq", "-q", "-k", "4", "-g", "u.fa", "-m", "1"],
                          {"u.fa": k3u, "r.fq": fq}))
    # K5: FASTA edge cases
    fa = (">multi\nAAATC\nGGATT\n>short\nAAAT\n>lower\naaatcggatt\n>withN\nAAATCNGATT\n>unmapped\nCCCCCCCCCC\n"
          ">nooverlap\nTTTTTTTTTTTT\n>bad char\nAAATCGGAXT\n>last\nAAATCGGATT")
    for m in (0, 1, 2):
        cases.append(cli_case("K5_fasta_edges_m%d" % m, ["-r", "r.fa", "-k", "4", "-g", "u.fa", "-m", str(m)],
                              {"u.fa": k3u, "r.fa": fa}))
    cases.append(cli_case("K5_fasta_edges_brute", ["-r", "r.fa", "-k", "4", "-g", "u.fa", "-m", "1", "-b"],
                          {"u.fa": k3u, "r.fa": fa}))
    # missing unitig file: empty index, everything "no overlap"
    cases.append(cli_case("missing_unitigs", ["-r", "r.fa", "-k", "4", "-g", "nope.fa"], {"r.fa": fa}))
    # unitig file with a short line in the middle: loading stops there (aligner.cpp:418)
    cases.append(cli_case("unitig_short_line_stops", ["-r", "r.fa", "-k", "4", "-g", "u.fa", "-m", "1"],
                          {"u.fa": ">1\nAAATC\n>2\nATC\n>3\nATCGGA\n>4\nGGATT\n", "r.fa": fa}))
    # a randomized mid-size case, two efforts
    import random
    rng = random.Random(1234)
    unitigs, haps = S.make_graph(rng, 7, genome_len=600, mean_unitig=15, n_variants=12)
    reads = S.make_reads(rng, haps, 300, 40, sub_rate=0.03, n_rate=0.005)
    ufa = "".join(">%d\n%s\n" % (i + 1, u) for i, u in enumerate(unitigs))
    rfa = "".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads))
    for e in (1, 2, 4):
        cases.append(cli_case("random_k7_e%d" % e, ["-r", "r.fa", "-k", "7", "-g", "u.fa", "-m", "2", "-e", str(e)],
                              {"u.fa": ufa, "r.fa": rfa}))
    return cases


def parse_case(name, content, k, fastq, chunk):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "reads")
        with open(p, "w", newline="") as f:
            f.write(content)
        recs = O.ref_harness_parse(p, k, fastq, chunk, d)
    return dict(name=name, content=content, k=k, fastq=int(fastq), chunk=chunk,
                records=[[h.decode("latin1"), r.decode("latin1")] for h, r in recs])


def parse_cases():
    cases = []
    fa = ">a\nACGTACGTAC\n>b\nACGT\nACGTAC\n\n>c\nAC\n>d\nACGTNACGTAC\n>e x y\nacgtacgtac\n>f\nACGTACGTAC"
    for chunk in (1, 2, 3, 10000):
        cases.append(parse_case("fasta_chunk%d" % chunk, fa, 4, False, chunk))
        cases.append(parse_case("fasta_nl_chunk%d" % chunk, fa + "\n", 4, False, chunk))
    cases.append(parse_case("fasta_blank_tail", fa + "\n\n\n", 4, False, 10000))
    cases.append(parse_case("fasta_crlf", fa.replace("\n", "\r\n"), 4, False, 10000))
    cases.append(parse_case("fasta_empty", "", 4, False, 10000))
    cases.append(parse_case("fasta_only_header", ">a\n", 4, False, 10000))
    cases.append(parse_case("fasta_len_eq_k", ">a\nACGT\n>b\nACGTA\n", 4, False, 10000))
    fq = "@a\nACGTACGTAC\n+\nIIIIIIIIII\n@b\nACG\n+\nIII\n@c\nAC\n+\nII\n@d\nACGTNACGTA\n+\nIIIIIIIIII\n@e\nACGTACGTAA\n+\nIIIIIIIIII"
    for chunk in (1, 2, 5, 10000):
        cases.append(parse_case("fastq_chunk%d" % chunk, fq, 4, True, chunk))
        cases.append(parse_case("fastq_nl_chunk%d" % chunk, fq + "\n", 4, True, chunk))
    cases.append(parse_case("fastq_empty", "", 4, True, 10000))
    cases.append(parse_case("fastq_truncated", "@a\nACGTACGTAC\n+\n", 4, True, 10000))
    return cases


def cli_correct_cases():
    """-c runs of the reference binary: corrected reads instead of paths (alignerGreedy.cpp:394-404)."""
    import random
    cases = []
    unitig_fa = open(os.path.join(REF_DIR, "unitig.fa")).read()
    reads_fa = open(os.path.join(REF_DIR, "reads.fa")).read()
    cases.append(cli_case("K1_c", ["-r", "reads.fa", "-k", "4", "-g", "unitig.fa", "-m", "2", "-c"],
                          {"unitig.fa": unitig_fa, "reads.fa": reads_fa}))
    k3u = ">1\nAAATC\n>2\nATCGGA\n>3\nGGATT\n"
    fa = (">multi\nAAATC\nGGATT\n>short\nAAAT\n>lower\naaatcggatt\n>withN\nAAATCNGATT\n>unmapped\nCCCCCCCCCC\n"
          ">nooverlap\nTTTTTTTTTTTT\n>bad char\nAAATCGGAXT\n>rc\nAATCCGATTT\n>last\nAAATCGGATT")
    for m in (0, 2):
        cases.append(cli_case("K5_c_m%d" % m, ["-r", "r.fa", "-k", "4", "-g", "u.fa", "-m", str(m), "-c"],
                              {"u.fa": k3u, "r.fa": fa}))
    fq = "@a\nAAATCGGATT\n+\nIIIIIIIIII\n@b\nAAATCGCATT\n+\nIIIIIIIIII\n"
    cases.append(cli_case("K4_fastq_c", ["-r", "r.fq", "-q", "-k", "4", "-g", "u.fa", "-m", "1", "-c"],
                          {"u.fa": k3u, "r.fq": fq}))
    for seed, k, rl in ((11, 7, 40), (12, 15, 60), (13, 31, 150)):
        rng = random.Random(seed)
        unitigs, haps = S.make_graph(rng, k, genome_len=2500, mean_unitig=max(3 * k, 30), n_variants=30)
        reads = S.make_reads(rng, haps, 300, rl, sub_rate=0.02, n_rate=0.002)
        ufa = "".join(">%d\n%s\n" % (i + 1, u) for i, u in enumerate(unitigs))
        rfa = "".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads))
        cases.append(cli_case("random_k%d_c" % k, ["-r", "r.fa", "-k", str(k), "-g", "u.fa", "-m", "2", "-c"],
                              {"u.fa": ufa, "r.fa": rfa}))
    return cases


def parse_fuzz_cases():
    """Fuzzed small texts through the reference's getReads (chunk 10000, what the CLI uses)."""
    import parse_gen as G
    cases = []
    for seed in range(150):
        for fastq in (False, True):
            cases.append(parse_case("fuzz_%s_%d" % ("fq" if fastq else "fa", seed), G.fuzz_text(seed, fastq), 4, fastq, 10000))
    return cases


def parse_big_cases():
    """Many-chunk files: only a digest of the reference's records is stored (the text is regenerated)."""
    import parse_gen as G
    cases = []
    for name in G.BIG_CASES:
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "reads")
            with open(p, "w", newline="") as f:
                f.write(G.big_text(name))
            recs = O.ref_harness_parse(p, 4, name.startswith("fq"), 10000, d)
        cases.append(dict(name=name, k=4, fastq=int(name.startswith("fq")), n_records=len(recs), sha256=G.digest(recs),
                          last=[[h.decode("latin1"), r.decode("latin1")] for h, r in recs[-2:]]))
    return cases


def main():
    if not O.have_ref():
        O.build(with_ref=True)
    if "--correct-extra" in sys.argv:  # only the -c fixtures
        json.dump(cli_correct_cases(), open(os.path.join(HERE, "cli_correct_cases.json"), "w"), indent=0, separators=(",", ":"))
        return
    if "--parse-extra" in sys.argv:  # only the parser fixtures added for the parallel reader
        json.dump(parse_fuzz_cases(), open(os.path.join(HERE, "parse_fuzz_cases.json"), "w"), indent=0, separators=(",", ":"))
        json.dump(parse_big_cases(), open(os.path.join(HERE, "parse_big_cases.json"), "w"), indent=0, separators=(",", ":"))
        return
    hc = known_answer_cases()
    for seed in range(40):
        c = S.fuzz_case(seed)
        hc.append(harness_case("fuzz_%d" % seed, c["k"], c["m"], c["effort"], c["unitigs"], c["reads"]))
    # one k=31 case with realistic read length
    import random
    rng = random.Random(31)
    unitigs, haps = S.make_graph(rng, 31, genome_len=4000, mean_unitig=60, n_variants=25)
    reads = S.make_reads(rng, haps, 150, 150, sub_rate=0.01, n_rate=0.001)
    for m in (0, 2, 5):
        hc.append(harness_case("k31_m%d" % m, 31, m, 2, unitigs, reads))
    json.dump(hc, open(os.path.join(HERE, "harness_cases.json"), "w"), indent=0, separators=(",", ":"))
    json.dump(cli_cases(), open(os.path.join(HERE, "cli_cases.json"), "w"), indent=0, separators=(",", ":"))
    json.dump(parse_cases(), open(os.path.join(HERE, "parse_cases.json"), "w"), indent=0, separators=(",", ":"))
    print("harness cases:", len(hc))


if __name__ == "__main__":
    main()
