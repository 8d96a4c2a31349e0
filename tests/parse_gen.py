"""Deterministic read-file texts for the parser tests (small fuzz texts and large many-chunk files)."""
import hashlib
import random


def fuzz_text(seed, fastq):
    """A short text made of everything a parser can trip over: consecutive '>' lines, blank lines,
    invalid characters, reads at/below the length limits, CRLF, missing / repeated final newline."""
    rng = random.Random(seed)
    lines = []
    for _ in range(rng.randrange(0, 40)):
        r = rng.random()
        if r < 0.30:
            lines.append((">" if not fastq or rng.random() < 0.3 else "@") + "h%d" % rng.randrange(100))
        elif r < 0.70:
            lines.append("".join(rng.choice("ACGTN" if rng.random() < 0.9 else "ACGTNacgtX") for _ in range(rng.randrange(0, 14))))
        elif r < 0.78:
            lines.append("")
        elif r < 0.86:
            lines.append("+")
        elif r < 0.92:
            lines.append(">")
        else:
            lines.append("IIII" + "#" * rng.randrange(0, 6))
    if fastq and rng.random() < 0.6:
        # well-formed 4-line records most of the time, so that valid reads actually occur
        lines = []
        for i in range(rng.randrange(0, 12)):
            s = "".join(rng.choice("ACGTN") for _ in range(rng.randrange(0, 12)))
            if rng.random() < 0.1:
                s += "x"
            lines += ["@r%d" % i, s, "+", "I" * len(s)]
        if lines and rng.random() < 0.3:
            lines = lines[:-rng.randrange(1, 4)]  # truncated last record
    nl = "\r\n" if rng.random() < 0.05 else "\n"
    text = nl.join(lines)
    text += nl * rng.choice([0, 1, 1, 1, 2, 3])
    return text


def big_text(name):
    """Many getReads chunks (10000 reads each): what the parallel reader cuts and the FASTQ tail rules
    (phantom duplicate of the last read iff the record count is not a multiple of 10000)."""
    kind, n, tail = name.split("_")
    n = int(n)
    rng = random.Random(hash_seed(name))
    out = []
    for i in range(n):
        ln = rng.choice([3, 5, 8, 12, 20, 33])
        s = "".join(rng.choice("ACGT") for _ in range(ln))
        u = rng.random()
        if u < 0.02:
            s = s[:2]  # too short
        elif u < 0.04:
            s = s[:-1] + "x"  # invalid character
        elif u < 0.06:
            s = s[:-1] + "N"
        if kind == "fq":
            out.append("@q%d\n%s\n+\n%s" % (i, s, "I" * len(s)))
        else:
            if rng.random() < 0.2 and len(s) > 4:  # multi-line record
                c = rng.randrange(1, len(s))
                s = s[:c] + "\n" + s[c:]
            out.append(">f%d some text\n%s" % (i, s))
    text = "\n".join(out)
    return text + {"nonl": "", "nl": "\n", "nl2": "\n\n"}[tail]


def hash_seed(name):
    return int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)


BIG_CASES = ["fq_20000_nl", "fq_20000_nonl", "fq_20001_nl", "fq_19999_nonl", "fq_30000_nl2", "fq_10000_nl", "fq_9999_nl",
             "fa_30000_nl", "fa_25000_nonl", "fa_25001_nl2"]


def digest(records):
    h = hashlib.sha256()
    for a, b in records:
        h.update(a if isinstance(a, bytes) else a.encode("latin1"))
        h.update(b"\t")
        h.update(b if isinstance(b, bytes) else b.encode("latin1"))
        h.update(b"\n")
    return h.hexdigest()
