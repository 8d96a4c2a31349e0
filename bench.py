#!/usr/bin/env python
"""bench.py -- reads/s of the read-to-de-Bruijn-graph mapping path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (pack + seed scan/index probe + greedy extension) over one batch
set of synthetic reads: BASELINE.json configs[1] -- simulated 5 Mbp diploid genome, its k=31 unitigs,
10 M 150 bp reads at 1 % substitution error, -m 2, greedy mode, effort 2 -- per GPU (weak scaling).

  value : whole-job reads/s with the reads already resident in HBM (device time, CUDA events on the
          launching stream, max over ranks)
  e2e   : the same metric through the C ABI with HOST (pinned) buffers: every step moves the reads host->device
          and the result records + paths device->host. Three forms of the public API are timed -- read
          characters over PCIe (bgr_slot_upload), reads packed to 2 bits per base by the host cores inside the
          timed region (bgr_pack_reads + bgr_slot_upload_packed), and a mix of both -- the fastest is the
          headline, all three are reported under e2e.forms
  roofline     : the seed-scan / index-probe kernel (k_scan_tiles_ws) against the measured HBM peak
  cpu_baseline : the reference `bgreat -t <all cores>` on a bounded sample of the same reads

`--impl reference` times the reference's own CPU implementation instead (oracle/_ref/bgreat, built from
the unmodified sources; falls back to the oracle port if that binary cannot run on this host).
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E2E_DEPTH = 4  # batches in flight on the end-to-end leg (slots 1..E2E_DEPTH; slot 0 holds the resident leg)
K = 31
ERRORS = 2
EFFORT = 2
READ_LEN = 150
SUB_RATE = 0.01
EXHAUSTIVE = False
WORKLOAD = "configs[1]: synthetic 5 Mbp diploid genome (SNP 1/150), k=31 unitigs, 150 bp reads at 1% substitutions, -m 2 -e 2, greedy"


ROOFLINE_NOTE = ("achieved = SURVEY 8(d) algorithmic bytes (one 32-byte index sector per (k-1)-mer window the reference "
                 "examines + 40 B of packed read) / measured kernel time. The kernel answers most windows from a Bloom "
                 "filter (shared memory for small indexes, L2 for large ones) and only verifies survivors in the key "
                 "table, so it moves far fewer DRAM bytes than that model (see traffic) and frac can exceed 1: with the "
                 "filter in shared memory it is bound by the integer ALU pipe (80 %% of peak, issue slots 68 %% busy: "
                 "profiles/r01_ncu_k_scan_tiles_v9_5Mbp.txt), with the filter in L2 by L2 sector throughput (one random "
                 "sector per window, 1.6 G sectors in 5.8 ms: profiles/r01_ncu_k_scan_tiles_v9_1Gbp.txt). "
                 "Index blob: %.1f MB.")


def configure(args):
    """The default run is BASELINE.json configs[1] (the configuration the metric is quoted on). The other
    configs (larger graphs, -m sweep, exhaustive mode) are selected with flags and named in `workload`."""
    global ERRORS, EFFORT, SUB_RATE, EXHAUSTIVE, WORKLOAD
    ERRORS, EFFORT, SUB_RATE, EXHAUSTIVE = args.errors, args.effort, args.sub_rate, args.mode == "exhaustive"
    default = (args.genome == 5_000_000 and ERRORS == 2 and EFFORT == 2 and SUB_RATE == 0.01 and not EXHAUSTIVE)
    if not default:
        g = args.genome
        gs = "%g Gbp" % (g / 1e9) if g >= 1e9 else "%g Mbp" % (g / 1e6)
        WORKLOAD = ("%ssynthetic %s diploid genome (SNP 1/150), k=31 unitigs, 150 bp reads at %g%% substitutions, "
                    "-m %d -e %d, %s" % ((args.label + ": ") if args.label else "", gs, 100 * SUB_RATE, ERRORS, EFFORT,
                                         args.mode))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--batch", type=int, default=1 << 20, help="reads per batch on the end-to-end leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the command-line leg (files in /dev/shm)")
    ap.add_argument("--cli-reads", type=int, default=10_000_000, help="reads of the command-line leg")
    ap.add_argument("--errors", type=int, default=2, help="-m")
    ap.add_argument("--effort", type=int, default=2, help="-e")
    ap.add_argument("--sub-rate", type=float, default=0.01, help="substitution error rate of the reads")
    ap.add_argument("--mode", default="greedy", choices=["greedy", "exhaustive"])
    ap.add_argument("--mix", type=float, default=0.5, help="end-to-end leg 'mixed': fraction of batches the host packs")
    ap.add_argument("--label", default="", help="name of the BASELINE.json config this run stands for")
    return ap.parse_args()


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen([exe, "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, l in self.lines:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw))}


# ------------------------------------------------------------------------------ CPU reference arm

def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def scratch_dir():
    for d in ("/dev/shm", tempfile.gettempdir()):
        if os.path.isdir(d) and os.access(d, os.W_OK):
            return tempfile.mkdtemp(prefix="bgr_bench_", dir=d)
    return tempfile.mkdtemp(prefix="bgr_bench_")


class CpuReference:
    """The reference's own CPU implementation of the path, on this box's host cores."""

    def __init__(self, graph, n_sample, seed_first=0):
        from oracle import oracle as O
        from tools import synth
        self.O = O
        self.dir = scratch_dir()
        self.n = n_sample
        self.cores = min(host_cores(), 255)  # -t is kept in an unsigned char (reference aligner.h:80)
        self.unitigs = os.path.join(self.dir, "unitigs.fa")
        self.reads = os.path.join(self.dir, "reads.fa")
        self.empty = os.path.join(self.dir, "empty.fa")
        graph.write_unitigs(self.unitigs)
        open(self.empty, "w").close()
        self.kind = "reference"
        self.index_s = None
        if not os.access(O.REF_BGREAT, os.X_OK) or not self._binary_runs():
            self.kind = "port"  # single-threaded C restatement: keep the sample small
            self.cores = 1
            self.n = n_sample = min(n_sample, 300_000)
        seqs, off = graph.make_reads(n_sample, READ_LEN, SUB_RATE, first=seed_first)
        synth.write_reads(self.reads, seqs, off, n_sample, first=seed_first)
        self.sample_seqs, self.sample_off = seqs, off

    def _run_binary(self, reads_file):
        t0 = time.perf_counter()
        out = subprocess.run([self.O.REF_BGREAT, "-r", reads_file, "-g", self.unitigs, "-k", str(K), "-m", str(ERRORS),
                              "-e", str(EFFORT), "-t", str(self.cores), "-f", os.path.join(self.dir, "paths.ref"),
                              "-a", os.path.join(self.dir, "na.ref")] + (["-b"] if EXHAUSTIVE else []), capture_output=True, text=True, cwd=self.dir,
                             timeout=3600)
        dt = time.perf_counter() - t0
        if out.returncode != 0:
            raise RuntimeError("reference bgreat failed: rc=%d %s" % (out.returncode, out.stderr[-200:]))
        reads = 0
        for l in out.stdout.splitlines():
            if l.startswith("Reads : "):
                reads = int(l.split(":")[1])
        return dt, reads

    def _binary_runs(self):
        try:
            # index-only run (empty read file): its wall time is subtracted from every step, which
            # matches the scope of the reference's own "Mapping in seconds" (aligner.cpp:551,594)
            a, _ = self._run_binary(self.empty)
            b, _ = self._run_binary(self.empty)
            self.index_s = min(a, b)
            return True
        except Exception as e:  # e.g. SIGILL on a host without AVX2
            sys.stderr.write("bench: reference binary unusable here (%s); using the oracle port\n" % e)
            return False

    def step(self):
        """-> (reads mapped, seconds of mapping)"""
        if self.kind == "reference":
            dt, reads = self._run_binary(self.reads)
            return reads, max(dt - self.index_s, 1e-6)
        o = self.O.Oracle(K, ERRORS, EFFORT)
        o.index_file(self.unitigs)
        t0 = time.perf_counter()
        o.run(self.O.EXHAUSTIVE if EXHAUSTIVE else self.O.GREEDY, self.reads, False, os.path.join(self.dir, "paths.port"), os.path.join(self.dir, "na.port"))
        return self.n, time.perf_counter() - t0

    def windows_per_read(self, n=20000):
        """W of SURVEY 8(d): (k-1)-mer windows the reference examines per read (incl. RC retries),
        from the oracle's instrumentation on a prefix of the sample."""
        o = self.O.Oracle(K, ERRORS, EFFORT)
        o.index_file(self.unitigs)
        b = self.sample_seqs.tobytes()
        n = min(n, self.n)
        for i in range(n):
            o.align(b[int(self.sample_off[i]):int(self.sample_off[i + 1])],
                    self.O.EXHAUSTIVE if EXHAUSTIVE else self.O.GREEDY)
        c = o.counters()
        return c["windows"] / float(n), c

    def describe(self):
        return "%d reads of the same synthetic set (%s), -t %d" % (self.n, "oracle/_ref/bgreat, index time subtracted"
                                                                  if self.kind == "reference" else "oracle C port",
                                                                  self.cores)

    def cleanup(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def cli_leg(graph, n_reads):
    """the drop-in `bgreat` command line end to end on page-cached files (/dev/shm): FASTA in,
    paths + notAligned.fa out, all host cores. Bounded sample; returns None if anything is missing."""
    from tools import synth
    exe = os.path.join(ROOT, "bgreat_b200", "bin", "bgreat")
    if not os.access(exe, os.X_OK):
        return None
    d = scratch_dir()
    try:
        graph.write_unitigs(os.path.join(d, "unitigs.fa"))
        seqs, off = graph.make_reads(n_reads, READ_LEN, SUB_RATE)
        synth.write_reads(os.path.join(d, "reads.fa"), seqs, off, n_reads)
        del seqs, off
        cores = min(host_cores(), 255)
        env = dict(os.environ, BGR_VERBOSE="1")
        t0 = time.perf_counter()
        p = subprocess.run([exe, "-r", "reads.fa", "-g", "unitigs.fa", "-k", str(K), "-m", str(ERRORS), "-e", str(EFFORT),
                            "-t", str(cores), "-f", "paths", "-a", "notAligned.fa"] + (["-b"] if EXHAUSTIVE else []),
                           cwd=d, capture_output=True, text=True, env=env, timeout=900)
        wall = time.perf_counter() - t0
        if p.returncode != 0:
            return None
        marks = {}
        for l in p.stderr.splitlines():  # "[bgreat] index built   t=0.698 s"
            if l.startswith("[bgreat]") and " t=" in l:
                marks[l[9:l.index(" t=")].strip()] = float(l.split(" t=")[1].split()[0])
        mapping = marks.get("mapping done", wall) - marks.get("index built", 0.0)
        out_bytes = sum(os.path.getsize(os.path.join(d, f)) for f in ("paths", "notAligned.fa"))
        return {"reads": n_reads, "threads": cores, "wall_s": wall, "mapping_s": mapping,
                "reads_per_s_mapping": n_reads / max(mapping, 1e-9), "reads_per_s_wall": n_reads / wall,
                "input_bytes": os.path.getsize(os.path.join(d, "reads.fa")), "output_bytes": out_bytes,
                "what": "bgreat_b200/bin/bgreat on FASTA in /dev/shm -> paths + notAligned.fa; mapping_s excludes "
                        "CUDA context creation and index build, like the reference's 'Mapping in seconds'"}
    except Exception as e:  # never let the optional leg break the bench line
        sys.stderr.write("bench: cli leg skipped (%s)\n" % e)
        return None
    finally:
        shutil.rmtree(d, ignore_errors=True)


def auto_cpu_sample(args):
    if args.cpu_sample:
        return min(args.cpu_sample, args.reads)
    # ~25 k reads/s/thread measured for the reference -> about 8 s of work
    return int(min(args.reads, 200_000 * min(host_cores(), 255)))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tools import synth
    graph = synth.Graph(args.genome, K, seed=1)
    n = auto_cpu_sample(args)
    ref = CpuReference(graph, n)
    try:
        for _ in range(args.warmup):
            ref.step()
        t, reads = 0.0, 0
        for _ in range(args.steps):
            r, dt = ref.step()
            reads += r
            t += dt
        v = reads / t
        line = {
            "impl": "reference", "metric": "reads/s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "char", "data": "synthetic",
            "config": {"workload": WORKLOAD, "genome_bp": args.genome, "k": K, "errors": ERRORS, "effort": EFFORT,
                       "read_len": READ_LEN, "reads_per_step": n},
            "bases_per_s": v * READ_LEN,
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": ref.cores, "kind": ref.kind,
                             "sample": ref.describe()},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
    finally:
        ref.cleanup()


# ------------------------------------------------------------------------------ B200 arm

def bind_to_gpu_numa_node(local):
    """Run this rank (and therefore allocate its pinned staging buffers, first touch) on the NUMA node its
    GPU hangs off: with 8 ranks streaming 50 GB/s each out of host memory, buffers on the wrong socket
    halve the end-to-end rate. Returns a description for the bench line; does nothing when the topology
    cannot be read. BGR_BENCH_NUMA=0 disables it."""
    if os.environ.get("BGR_BENCH_NUMA", "1") == "0":
        return "off"
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        dom, rest = out.split(":", 1)
        dev = "%s:%s" % (dom[-4:], rest)  # sysfs uses a 4-digit domain
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % dev).read())
        if node < 0:
            return "no numa information"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return "node %d has no usable cpu" % node
        os.sched_setaffinity(0, cpus)
        return "node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:
        return "unavailable (%s)" % type(e).__name__


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from tools import synth
    import bgreat_b200
    from bgreat_b200 import _abi as A

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the mapping path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L = A.lib()
    n_reads = args.reads

    # ---- graph + index: built once on rank 0, replicated by one NCCL broadcast of the flat blob
    graph = synth.Graph(args.genome, K, seed=1)
    al = bgreat_b200.Aligner(K, ERRORS, EFFORT, device=local, n_slots=1 + E2E_DEPTH)
    from bgreat_b200 import shard
    t0 = time.perf_counter()

    def build_index():
        useqs, uoff = graph.unitig_arrays()
        al.index_unitigs_packed(useqs, uoff)

    shard.replicate_index(al, build_index, src=0)  # rank 0 builds; NCCL broadcast of the blob; import
    index_build_s = time.perf_counter() - t0
    stats = al.index_stats()

    # ---- this rank's reads (weak scaling: reads [rank*n, (rank+1)*n)), in pinned host memory
    seq_t = torch.empty(n_reads * READ_LEN + 64, dtype=torch.uint8, pin_memory=True)
    off_t = torch.empty(n_reads + 1, dtype=torch.int64, pin_memory=True)
    seqs = seq_t.numpy()
    off = off_t.numpy().view(np.uint64)
    graph.make_reads(n_reads, READ_LEN, SUB_RATE, first=rank * n_reads, seqs=seqs, off=off)

    params = al.params(A.MODE_EXHAUSTIVE if EXHAUSTIVE else A.MODE_GREEDY)
    ctx = al.ctx

    def ck(rc):
        if rc:
            raise A.BgrError(rc, L.bgr_last_error(ctx).decode())

    clocks = ClockSampler(local)
    clocks.start()

    # ================= leg 1: reads resident in HBM (slot 0 holds the whole step) =================
    ck(L.bgr_slot_upload(ctx, 0, seqs.ctypes.data, off.ctypes.data, n_reads))
    ck(L.bgr_slot_wait(ctx, 0, None))
    ms = (C.c_float * 3)()

    def resident_step():
        ck(L.bgr_slot_run(ctx, 0, C.byref(params)))
        ck(L.bgr_slot_wait(ctx, 0, None))
        ck(L.bgr_slot_kernel_ms(ctx, 0, ms))
        return ms[0], ms[1], ms[2]

    for _ in range(args.warmup):
        resident_step()
    al.reset_counters()
    barrier()
    launches0 = al.launch_count()
    t_res0 = time.perf_counter()
    dev = np.zeros(3)
    for _ in range(args.steps):
        dev += np.array(resident_step())
    barrier()
    t_res1 = time.perf_counter()
    launches = al.launch_count() - launches0
    counters = shard.reduce_counters(al.counters(), device="cuda")  # whole-job totals
    dev_ms = torch.tensor(dev, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    dev_ms = dev_ms.cpu().numpy()
    steps = max(args.steps, 1)
    value = world * n_reads * steps / (dev_ms[2] / 1e3)

    # the same leg with the reads resident as the 2-bit stream the packed / mixed end-to-end forms upload
    # (extra information: `value` above stays on the character form)
    pk_all = torch.empty(n_reads * READ_LEN // 32 + 16, dtype=torch.int64, pin_memory=True)
    bad_all = torch.empty(n_reads // 32 + 2, dtype=torch.int32, pin_memory=True)
    nb = C.c_uint32()
    ck(L.bgr_pack_reads(seqs.ctypes.data, off.ctypes.data, n_reads, max(1, host_cores() // max(1, world)),
                        pk_all.data_ptr(), bad_all.data_ptr(), C.byref(nb)))
    ck(L.bgr_slot_upload_packed(ctx, 0, pk_all.data_ptr(), off.ctypes.data, n_reads, bad_all.data_ptr(), nb.value,
                                seqs.ctypes.data))
    ck(L.bgr_slot_wait(ctx, 0, None))
    for _ in range(args.warmup):
        resident_step()
    barrier()
    dev_pk = np.zeros(3)
    for _ in range(args.steps):
        dev_pk += np.array(resident_step())
    barrier()
    dev_pk_t = torch.tensor(dev_pk, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev_pk_t, op=dist.ReduceOp.MAX)
    dev_pk = dev_pk_t.cpu().numpy()
    del pk_all, bad_all

    # ================= leg 2: end to end through the C ABI with host buffers =================
    B = min(args.batch, n_reads)
    n_batches = (n_reads + B - 1) // B
    res_t = torch.empty((n_reads, 2), dtype=torch.int32, pin_memory=True)
    pool_cap = B * 16
    pool_t = [torch.empty(pool_cap, dtype=torch.int32, pin_memory=True) for _ in range(E2E_DEPTH)]
    boff = [np.ascontiguousarray(off[b * B:min(n_reads, (b + 1) * B) + 1] - off[b * B]) for b in range(n_batches)]
    boff_t = []
    for a in boff:  # per-batch offsets (each batch starts at 0) also pinned
        t = torch.empty(len(a), dtype=torch.int64, pin_memory=True)
        t.numpy().view(np.uint64)[:] = a
        boff_t.append(t)
    res_np = res_t.numpy()
    d2h = [0]
    # packed upload: the host packs every batch to 2 bits per base inside the timed region (bgr_pack_reads on
    # all host cores of this rank), the 2-bit stream is what crosses PCIe
    pack_threads = max(1, min(host_cores() // max(1, world), 64))
    pk_t = [torch.empty(B * READ_LEN // 32 + 16, dtype=torch.int64, pin_memory=True) for _ in range(E2E_DEPTH)]
    bad_t = [torch.empty(B // 32 + 2, dtype=torch.int32, pin_memory=True) for _ in range(E2E_DEPTH)]
    h2d_packed = [0]

    def e2e_step(form="characters"):
        """form: 'characters' | 'packed' | 'mixed' (a fraction args.mix of the batches is packed by the host
        cores while the others cross PCIe as characters: the two forms draw on different resources)"""
        frac = {"characters": 0.0, "packed": 1.0, "mixed": args.mix}[form]
        used = C.c_uint64()
        n_bad = C.c_uint32()
        inflight = []
        d2h[0] = 0
        h2d_packed[0] = 0
        for b in range(n_batches + 1):
            if b < n_batches:
                slot = 1 + (b % E2E_DEPTH)
                q = b % E2E_DEPTH
                packed = int((b + 1) * frac) > int(b * frac)
                lo = b * B
                n = min(n_reads, lo + B) - lo
                src = seqs.ctypes.data + int(off[lo])
                if packed:
                    ck(L.bgr_pack_reads(src, boff_t[b].data_ptr(), n, pack_threads, pk_t[q].data_ptr(),
                                        bad_t[q].data_ptr(), C.byref(n_bad)))
                    ck(L.bgr_slot_upload_packed(ctx, slot, pk_t[q].data_ptr(), boff_t[b].data_ptr(), n,
                                                bad_t[q].data_ptr(), n_bad.value, src))
                    h2d_packed[0] += ((int(boff[b][-1]) + 31) // 32) * 8 + (n + 1) * 8 + ((n + 31) // 32) * 4
                else:
                    ck(L.bgr_slot_upload(ctx, slot, src, boff_t[b].data_ptr(), n))
                    h2d_packed[0] += int(boff[b][-1]) + (n + 1) * 8
                ck(L.bgr_slot_run(ctx, slot, C.byref(params)))
                ck(L.bgr_slot_download(ctx, slot, res_np.ctypes.data + lo * 8, pool_t[q].data_ptr(), pool_cap))
                inflight.append((slot, n))
            if len(inflight) == E2E_DEPTH or (b == n_batches and inflight):
                slot, n = inflight.pop(0)
                ck(L.bgr_slot_wait(ctx, slot, C.byref(used)))
                d2h[0] += n * 8 + used.value * 4
        while inflight:
            slot, n = inflight.pop(0)
            ck(L.bgr_slot_wait(ctx, slot, C.byref(used)))
            d2h[0] += n * 8 + used.value * 4

    def time_e2e(form):
        for _ in range(max(3, args.warmup)):
            e2e_step(form)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step(form)
        barrier()
        t1 = time.perf_counter()
        sec = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return float(sec.item()), t1, int((res_np[:, 0] & 1).sum()), d2h[0], h2d_packed[0]

    # both upload forms of the public API are timed; the headline is the faster one, the other is reported beside it
    what = {"characters": "bgr_slot_upload: read characters + offsets cross PCIe (8 bits per base)",
            "packed": "bgr_pack_reads (host cores, inside the timed region) + bgr_slot_upload_packed: 2 bits per base "
                      "+ offsets + flag bitmap cross PCIe",
            "mixed": "a fraction %g of the batches is packed by the host cores (as in 'packed') while the others "
                     "cross PCIe as characters: the two forms load different resources" % args.mix}
    e2e_forms, t_e1, mapped = {}, 0.0, set()
    for form in ("characters", "packed", "mixed"):
        sec, t_e1, m, d2h_f, h2d_f = time_e2e(form)
        mapped.add(m)
        e2e_forms[form] = {"value": world * n_reads * steps / sec, "unit": "reads/s", "ms_per_step": 1e3 * sec / steps,
                           "h2d_bytes_per_step": int(h2d_f), "d2h_bytes_per_step": int(d2h_f), "what": what[form]}
        if form != "characters":
            e2e_forms[form]["host_pack_threads"] = pack_threads
    e2e_best = max(e2e_forms, key=lambda f: e2e_forms[f]["value"])
    e2e_s_best = e2e_forms[e2e_best]["ms_per_step"] * steps / 1e3
    e2e_value = e2e_forms[e2e_best]["value"]
    clocks.stop()
    clk = clocks.summary(t_res0, t_e1)
    # sanity: every end-to-end leg reproduces the same outcome
    mapped_e2e = mapped.pop()
    if mapped:
        raise SystemExit("bench.py: the upload forms disagree on the number of mapped reads")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ================= roofline of the probe kernel + CPU baseline (rank 0) =================
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            # measured DRAM bytes of one 10 M-read launch of the scan kernel on a graph of this size
            # (ncu --set full, tools/gpu_profile.sh), scaled to this run's reads per launch
            t10 = json.load(open(prof)).get("k_scan_tiles@%d" % args.genome)
            traffic = None if t10 is None else t10 * n_reads / 1e7
        except Exception:
            traffic = None

    cpu = None
    w_per_read, w_src = 121.5, "SURVEY.md 8(d) figure"
    if not args.no_cpu_baseline:
        ref = CpuReference(graph, auto_cpu_sample(args))
        try:
            if world == 1:
                r, dt = ref.step()
                cpu = {"value": r / dt, "unit": "reads/s", "cores": ref.cores, "kind": ref.kind,
                       "sample": ref.describe()}
            w_per_read, _ = ref.windows_per_read()
            w_src = "oracle instrumentation, 20000 reads of this set"
        finally:
            ref.cleanup()
    algo_bytes = n_reads * (w_per_read * 32.0 + 40.0)
    scan_ms = dev_ms[0] / steps
    achieved = algo_bytes / (scan_ms / 1e3) / 1e9
    line = {
        "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms[2] / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "genome_bp": args.genome, "k": K, "errors": ERRORS, "effort": EFFORT,
                   "read_len": READ_LEN, "reads_per_gpu_per_step": n_reads, "unitigs": stats["unitigs"],
                   "border_keys": stats["keys"], "index_bytes": stats["device_bytes"],
                   "l2": "inputs (1.5 GB of read characters per step) exceed L2; no flush needed",
                   "e2e_batch_reads": B, "parallelism": "reads sharded across GPUs, index replicated",
                   "rank0_numa_binding": numa},
        "bases_per_s": value * READ_LEN,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": e2e_forms[e2e_best]["h2d_bytes_per_step"],
                "d2h_bytes_per_step": e2e_forms[e2e_best]["d2h_bytes_per_step"], "ms_per_step": 1e3 * e2e_s_best / steps,
                "upload": e2e_best, "forms": e2e_forms},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_scan_tiles_ws (pack + seed scan / index probe)", "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                     "windows_per_read": w_per_read, "windows_source": w_src, "kernel_ms": scan_ms,
                     "note": ROOFLINE_NOTE % (stats["device_bytes"] / 1e6)},
        "kernel_ms": {"pack_scan": scan_ms, "extend": dev_ms[1] / steps},
        "packed_input": {"value": world * n_reads * steps / (dev_pk[2] / 1e3), "unit": "reads/s",
                         "kernel_ms": {"scan": dev_pk[0] / steps, "extend": dev_pk[1] / steps},
                         "what": "same resident leg with the reads in HBM as the 2-bit stream of a packed upload"},
        "wall_ms_per_step_resident": 1e3 * (t_res1 - t_res0) / steps,
        "outcome": {"scope": "all ranks, per step", "reads": counters["reads"] // steps, "aligned": counters["aligned"] // steps,
                    "not_aligned": counters["not_aligned"] // steps, "no_overlap": counters["no_overlap"] // steps,
                    "mapped_e2e_last_step": mapped_e2e},
        "index_build_s": index_build_s,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if world == 1 and not args.no_cli:
        al.close()  # the command line creates its own context
        line["cli"] = cli_leg(graph, min(n_reads, args.cli_reads))
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def quiet_stdout():
    """Libraries (NCCL's version banner, for one) write to file descriptor 1; the contract is ONE JSON line on
    stdout. Everything else is sent to stderr; the returned file object is the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse_args()
    configure(args)
    out = quiet_stdout()
    sys.stdout = out
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
