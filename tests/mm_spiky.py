"""Sequences whose minimizer (bgr_common.h: mm_hash over the 15 16-mers of a 30-mer) changes at nearly every window:
reads cut from them have chunks of 32 windows with far more runs than the staged seed scan has shared-memory slots
(MM_SLOTS = 10, bgr_scan.cuh), i.e. they take the path that redoes a chunk with the loads in place. Test input only;
the checker stays the oracle."""
import random

M32 = 0xFFFFFFFF
CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def mm_hash(s, r):
    return (((s * r) & M32) * 0x9E3779B1 + s * 0x85EBCA77 + r * 0x85EBCA77) & M32


def code16(seq):
    c = 0
    for ch in seq:
        c = (c << 2) | CODE[ch]
    return c


def rc16(c):
    r = 0
    for _ in range(16):
        r = (r << 2) | (3 - (c & 3))
        c >>= 2
    return r


def hash16(seq):
    c = code16(seq)
    return mm_hash(c, rc16(c))


def minimizers(seq):
    """smallest 16-mer hash of every 30-mer of seq"""
    h = [hash16(seq[i:i + 16]) for i in range(len(seq) - 15)]
    return [min(h[i:i + 15]) for i in range(len(seq) - 29)]


def runs_in(mins, lo, hi):
    """windows of [lo, hi) whose minimizer differs from the window before (the first always counts)"""
    return 1 + sum(1 for j in range(lo + 1, hi) if mins[j] != mins[j - 1])


def spiky_sequence(rng, n, beam=64):
    """Beam search, base by base, for a sequence with as many minimizer changes as possible. A state is the last
    fifteen 16-mer hashes (the youngest fourteen stay in the next window), the code of the youngest 16-mer and of its
    reverse complement, the bases so far and the number of changes so far."""
    start = [rng.choice("ACGT") for _ in range(16)]
    c = code16(start)
    states = [(0, [mm_hash(c, rc16(c))], c, rc16(c), start)]
    for _ in range(n - 16):
        nxt = []
        for score, h, c, r, seq in states:
            old_min = min(h) if len(h) == 15 else None
            keep = h[-14:]
            for b in range(4):
                c2 = ((c << 2) | b) & M32
                r2 = (r >> 2) | ((3 - b) << 30)
                x = mm_hash(c2, r2)
                h2 = keep + [x]
                changed = len(h2) == 15 and old_min is not None and min(h2) != old_min
                nxt.append((score + (1 if changed else 0), h2, c2, r2, seq + ["ACGT"[b]]))
        rng.shuffle(nxt)
        nxt.sort(key=lambda t: -t[0])
        states = nxt[:beam]
    return "".join(states[0][4])


if __name__ == "__main__":
    rng = random.Random(5)
    s = spiky_sequence(rng, 3000)
    m = minimizers(s)
    per_chunk = [runs_in(m, j, j + 32) for j in range(0, len(m) - 32, 7)]
    print("windows", len(m), "mean runs per 32 windows", sum(per_chunk) / len(per_chunk), "min", min(per_chunk),
          "share above 10:", sum(1 for x in per_chunk if x > 10) / len(per_chunk))
    r = "".join(rng.choice("ACGT") for _ in range(3000))
    m = minimizers(r)
    per_chunk = [runs_in(m, j, j + 32) for j in range(0, len(m) - 32, 7)]
    print("random: mean", sum(per_chunk) / len(per_chunk), "max", max(per_chunk))
