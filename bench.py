#!/usr/bin/env python
"""bench.py -- reads/s of the read-to-de-Bruijn-graph mapping path on B200 (BASELINE.json metric).

Default workload = BASELINE.json configs[4], one GPU's shard: synthetic 3 Gbp diploid genome (SNP 1/150, 0.1 % of the
sequence in 2..6-copy repeat families), its k=31 unitigs (60 M unitigs, 40 M border keys, 13 GB index replicated in
HBM), 150 bp reads at 1 % substitution error, -m 2 -e 2, greedy mode, 10 M reads per GPU and pass (weak scaling).
A "step" is PASSES passes of the hot path (pack + seed scan / index probe + greedy extension + path compaction) over
the batch, so that the timed region of the default run lasts more than a second. The other configs are selected
with --genome / --errors / --mode / --sub-rate and are named in `config.workload`.

  value : whole-job reads/s with the reads already resident in HBM (device time, CUDA events on the launching
          stream, max over ranks)
  e2e   : the same metric through the C ABI with HOST (pinned) buffers, every pass moving the reads host->device
          and the result records + paths device->host. The headline is the form the reference's own API has --
          read CHARACTERS (bgr_slot_upload); the two other upload forms of the public API (reads packed to 2 bits
          per base by the host cores inside the timed region, and a mix of both) are reported beside it
  roofline     : the kernel that takes the most device time, and under `kernels` both hot kernels, each with the
                 bytes this design has to move (inputs once, outputs once), the DRAM bytes ncu measured for this
                 graph size, the resource that binds it, and -- separately, as model_frac -- SURVEY 8(d)'s model
  cpu_baseline : the reference `bgreat -t <all cores>` on a bounded sample of the same reads

`--impl reference` times the reference's own CPU implementation instead (oracle/_ref/bgreat, built from the
unmodified sources; falls back to the oracle port if that binary cannot run on this host): ONE process that
indexes the graph once and maps warm-up + steps copies of the sample file (`-r f,f,...,f`, aligner.cpp:554-586);
a step's time is the time between two file names appearing on its stdout.
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
REPEAT_FRAC = 0.001
PASSES = 8
DEFAULT_GENOME = 3_000_000_000
WORKLOAD = ""


def configure(args):
    """The default run is BASELINE.json configs[4] (the 3 Gbp graph the metric's HBM target is defined on), one
    10 M-read shard per GPU and pass. The other configs (smaller graphs, -m sweep, exhaustive mode) are selected
    with flags and named in `workload`."""
    global ERRORS, EFFORT, SUB_RATE, EXHAUSTIVE, WORKLOAD, REPEAT_FRAC, PASSES
    ERRORS, EFFORT, SUB_RATE, EXHAUSTIVE = args.errors, args.effort, args.sub_rate, args.mode == "exhaustive"
    REPEAT_FRAC, PASSES = args.repeat_frac, max(1, args.passes)
    g = args.genome
    gs = "%g Gbp" % (g / 1e9) if g >= 1e9 else "%g Mbp" % (g / 1e6)
    default = (g == DEFAULT_GENOME and ERRORS == 2 and EFFORT == 2 and SUB_RATE == 0.01 and not EXHAUSTIVE)
    label = args.label or ("configs[4], one shard of the 1 B reads per GPU" if default else "")
    WORKLOAD = ("%ssynthetic %s diploid genome (SNP 1/150, %g%% repeats), k=31 unitigs, 150 bp reads at %g%% "
                "substitutions, -m %d -e %d, %s" % ((label + ": ") if label else "", gs, 100 * REPEAT_FRAC, 100 * SUB_RATE,
                                                    ERRORS, EFFORT, args.mode))


def workload_config(args):
    """`config` of the JSON line: the workload, the same in both arms."""
    return {"workload": WORKLOAD, "genome_bp": args.genome, "repeat_frac": REPEAT_FRAC, "k": K, "errors": ERRORS,
            "effort": EFFORT, "read_len": READ_LEN, "sub_rate": SUB_RATE, "mode": args.mode,
            "reads_per_gpu_per_pass": args.reads, "passes_per_step": PASSES,
            "parallelism": "reads sharded across GPUs, index replicated",
            "l2": "inputs (1.5 GB of read characters per pass) exceed L2; no flush needed"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per pass")
    ap.add_argument("--passes", type=int, default=8, help="passes over the batch per step")
    ap.add_argument("--genome", type=int, default=DEFAULT_GENOME)
    ap.add_argument("--repeat-frac", type=float, default=0.001, help="fraction of the genome in repeat families")
    ap.add_argument("--batch", type=int, default=1 << 20, help="reads per batch on the end-to-end leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the command-line leg (files in /dev/shm)")
    ap.add_argument("--cli-reads", type=int, default=40_000_000, help="reads of the command-line leg")
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
    """The reference's own CPU implementation of the path, on this box's host cores: ONE `bgreat` process that
    indexes the graph once and then maps the sample file as many times as asked (-r f,f,...,f). The reference
    prints each read file's name when it opens it and "The End" after the last one (aligner.cpp:556-588), so the
    time between two of those lines is the mapping time of one file: no index-only run to subtract."""

    def __init__(self, graph, n_sample, seed_first=0):
        from oracle import oracle as O
        from tools import synth
        self.O = O
        self.dir = scratch_dir()
        self.n = n_sample
        self.cores = min(host_cores(), 255)  # -t is kept in an unsigned char (reference aligner.h:80)
        self.unitigs = os.path.join(self.dir, "unitigs.fa")
        self.reads = os.path.join(self.dir, "reads.fa")
        graph.write_unitigs(self.unitigs)
        self.kind = "reference"
        self.index_s = None
        if not os.access(O.REF_BGREAT, os.X_OK) or not self._binary_runs():
            self.kind = "port"  # single-threaded C restatement: keep the sample small
            self.cores = 1
            self.n = n_sample = min(n_sample, 300_000)
        seqs, off = graph.make_reads(n_sample, READ_LEN, SUB_RATE, first=seed_first)
        synth.write_reads(self.reads, seqs, off, n_sample, first=seed_first)
        del seqs, off

    def _binary_runs(self):
        try:  # e.g. SIGILL on a host without AVX2
            return subprocess.run([self.O.REF_BGREAT], capture_output=True, timeout=60).returncode == 0
        except Exception as e:
            sys.stderr.write("bench: reference binary unusable here (%s); using the oracle port\n" % e)
            return False

    def run(self, n_files):
        """-> per-file mapping seconds (n_files of them), reads per file"""
        if self.kind == "port":
            o = self.O.Oracle(K, ERRORS, EFFORT)
            o.index_file(self.unitigs)
            out = []
            for _ in range(n_files):
                t0 = time.perf_counter()
                o.run(self.O.EXHAUSTIVE if EXHAUSTIVE else self.O.GREEDY, self.reads, False,
                      os.path.join(self.dir, "paths.port"), os.path.join(self.dir, "na.port"))
                out.append(time.perf_counter() - t0)
            return out, self.n
        cmd = [self.O.REF_BGREAT, "-r", ",".join(["reads.fa"] * n_files), "-g", "unitigs.fa", "-k", str(K), "-m", str(ERRORS),
               "-e", str(EFFORT), "-t", str(self.cores), "-f", "paths.ref", "-a", "na.ref"] + (["-b"] if EXHAUSTIVE else [])
        t0 = time.perf_counter()
        p = subprocess.Popen(cmd, cwd=self.dir, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        marks, reads = [], 0
        for line in p.stdout:
            now = time.perf_counter()
            line = line.strip()
            if line == "reads.fa" or line == "The End":
                marks.append(now)
            elif line.startswith("Reads : "):
                reads = int(line.split(":")[1])
        rc = p.wait()
        if rc != 0 or len(marks) != n_files + 1:
            raise RuntimeError("reference bgreat failed: rc=%d, %d file marks" % (rc, len(marks)))
        self.index_s = marks[0] - t0
        return [marks[i + 1] - marks[i] for i in range(n_files)], reads // n_files

    def describe(self):
        return "%d reads of the same synthetic set (%s), -t %d; one process, index built once (%s s), time per file " \
               "from its stdout" % (self.n, "oracle/_ref/bgreat" if self.kind == "reference" else "oracle C port",
                                    self.cores, "%.0f" % self.index_s if self.index_s else "-")

    def cleanup(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def reference_model(n=20000):
    """What the reference does per read on this workload -- windows it examines, index queries, comparisons and
    compared bases -- from the oracle's instrumentation (these figures come from the spec model, not
    from the GPU implementation). They are properties of the reads and of the graph's local shape, not of the
    genome's length, so they are taken on a 5 Mbp graph of the same generator (the oracle would need minutes and
    tens of GB to index the 3 Gbp one)."""
    from oracle import oracle as O
    from tools import synth
    g = synth.Graph(5_000_000, K, seed=1, repeat_frac=REPEAT_FRAC)
    o = O.Oracle(K, ERRORS, EFFORT)
    o.index_unitigs(g.unitig_list())
    seqs, off = g.make_reads(n, READ_LEN, SUB_RATE)
    b = seqs.tobytes()
    for i in range(n):
        o.align(b[int(off[i]):int(off[i + 1])], O.EXHAUSTIVE if EXHAUSTIVE else O.GREEDY)
    c = o.counters()
    return {k_: c[k_] / float(n) for k_ in ("windows", "queries", "compares", "cmp_bases")}


def cpu_cache_path(args):
    key = "%d_%g_%d_%d_%g_%s_%d" % (args.genome, REPEAT_FRAC, ERRORS, EFFORT, SUB_RATE, args.mode, min(host_cores(), 255))
    return os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(), "bgr_cpu_baseline_%s.json" % key)


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
    # ~25 k reads/s/thread measured for the reference on the small graphs, about half that on the 3 Gbp one
    #  -> a few seconds of work per file
    return int(min(args.reads, 100_000 * min(host_cores(), 255)))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tools import synth
    graph = synth.Graph(args.genome, K, seed=1, repeat_frac=REPEAT_FRAC)
    n = auto_cpu_sample(args)
    ref = CpuReference(graph, n)
    del graph
    try:
        secs, per_file = ref.run(args.warmup + args.steps)
        timed = secs[args.warmup:]
        v = per_file * len(timed) / sum(timed)
        cpu = {"value": v, "unit": "reads/s", "cores": ref.cores, "kind": ref.kind, "sample": ref.describe()}
        line = {
            "impl": "reference", "metric": "reads/s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(timed) / max(len(timed), 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "char", "data": "synthetic",
            "config": workload_config(args), "bases_per_s": v * READ_LEN, "reads_per_step": per_file,
            "index_s": ref.index_s, "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        try:  # the B200 arm that follows on this box reports this measurement instead of indexing the graph again
            json.dump(dict(cpu, when=time.time()), open(cpu_cache_path(args), "w"))
        except Exception:
            pass
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
    # only the rank that builds the index needs the unitigs; the others regenerate the stretches of genome their
    # reads come from (tools/bgr_synth.cpp: nothing of genome size is stored)
    graph = synth.Graph(args.genome, K, seed=1, repeat_frac=REPEAT_FRAC, unitigs=(rank == 0))
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
        tot = np.zeros(3)
        for _ in range(PASSES):
            ck(L.bgr_slot_run(ctx, 0, C.byref(params)))
            ck(L.bgr_slot_wait(ctx, 0, None))
            ck(L.bgr_slot_kernel_ms(ctx, 0, ms))
            tot += (ms[0], ms[1], ms[2])
        return tot

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
    value = world * n_reads * PASSES * steps / (dev_ms[2] / 1e3)

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
        for _ in range(args.steps * PASSES):
            e2e_step(form)
        barrier()
        t1 = time.perf_counter()
        sec = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return float(sec.item()), t1, int((res_np[:, 0] & 1).sum()), d2h[0], h2d_packed[0]

    # all upload forms of the public API are timed; the headline is the character form
    what = {"characters": "bgr_slot_upload: read characters + offsets cross PCIe (8 bits per base)",
            "packed": "bgr_pack_reads (host cores, inside the timed region) + bgr_slot_upload_packed: 2 bits per base "
                      "+ offsets + flag bitmap cross PCIe",
            "mixed": "a fraction %g of the batches is packed by the host cores (as in 'packed') while the others "
                     "cross PCIe as characters: the two forms load different resources" % args.mix}
    e2e_forms, t_e1, mapped = {}, 0.0, set()
    for form in ("characters", "packed", "mixed"):
        sec, t_e1, m, d2h_f, h2d_f = time_e2e(form)
        mapped.add(m)
        e2e_forms[form] = {"value": world * n_reads * PASSES * steps / sec, "unit": "reads/s", "ms_per_step": 1e3 * sec / steps,
                           "h2d_bytes_per_step": int(h2d_f) * PASSES, "d2h_bytes_per_step": int(d2h_f) * PASSES,
                           "what": what[form]}
        if form != "characters":
            e2e_forms[form]["host_pack_threads"] = pack_threads
    e2e_best = "characters"  # the headline is the reference API's own shape; the other forms are reported beside it
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

    # ================= rooflines of the two hot kernels + CPU baseline (rank 0) =================
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        pass

    def measured_traffic(kernel):
        """DRAM bytes of one 10 M-read launch of `kernel` on a graph of this size (ncu --set full,
        tools/gpu_profile.sh -> profiles/roofline_traffic.json), scaled to this run's reads per launch"""
        t10 = prof.get("%s@%d" % (kernel, args.genome))
        return None if t10 is None else t10 * n_reads / 1e7

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cache = cpu_cache_path(args)
        try:
            c = json.load(open(cache))
            if time.time() - c.pop("when") < 3 * 3600:
                cpu = c
                cpu["sample"] += "; measured by `bench.py --impl reference` on this box %d min before this run" % (
                    (time.time() - os.path.getmtime(cache)) / 60)
        except Exception:
            cpu = None
        if cpu is None:
            ref = CpuReference(graph, auto_cpu_sample(args))
            try:
                secs, per_file = ref.run(3)  # one warm-up file, two timed
                cpu = {"value": 2 * per_file / sum(secs[1:]), "unit": "reads/s", "cores": ref.cores, "kind": ref.kind,
                       "sample": ref.describe()}
            finally:
                ref.cleanup()
    model = {"windows": 119.1, "queries": 2.853, "compares": 3.747, "cmp_bases": 155.1}
    model_src = "SURVEY.md 8(d) / round-1 oracle instrumentation"
    if not args.no_cpu_baseline:
        try:
            model = reference_model()
            model_src = "oracle instrumentation, 20000 reads on a 5 Mbp graph of the same generator"
        except Exception as e:
            sys.stderr.write("bench: oracle instrumentation skipped (%s)\n" % e)

    per_pass = steps * PASSES
    scan_ms, ext_ms = dev_ms[0] / per_pass, dev_ms[1] / per_pass
    listed = counters["aligned"] + counters["not_aligned"]  # reads the extension kernel looked at (greedy mode)
    listed_per_read = listed / max(counters["reads"], 1)
    verified_per_read = counters["verified"] / max(counters["reads"], 1)
    # ---- the bytes this design has to move per read, inputs once and outputs once
    #   scan: characters + offset in; packed words (40 B) + hit mask (16 B) + info (4 B) + result or list entry (8 B)
    #         out; one 8-byte filter block per run of windows sharing a minimizer (~W/8; one 4-byte word per window
    #         on indexes with the word filters... counted as W/8 blocks of 8 B here only for the minimizer tier) and
    #         one 32-byte key bucket per window that passes the filter (counter `verified`)
    #   extension (per read the scan lists): list entry, offsets, info, hit mask, packed words in (4+16+4+16+40);
    #         per seed a key bucket (32 B); per index query one entry side (64 B); per comparison the unitig words
    #         under it (cmp_bases/4 bytes, at least one 32-byte sector); path ints + result record out
    large = stats["keys"] > 131072
    filt = (model["windows"] / 8.0) * 8.0 if large else 0.0
    scan_bytes = READ_LEN + 8 + 40 + 16 + 4 + 8 + filt + 32.0 * verified_per_read
    lm = 1.0 / max(listed_per_read, 1e-9)  # the model's per-read figures are over all reads: per listed read
    cmp_b = max(32.0, model["cmp_bases"] / max(model["compares"], 1e-9) / 4.0)
    ext_bytes = listed_per_read * (80 + 32 * 1.15 + lm * (64.0 * model["queries"] + cmp_b * model["compares"]) + 4 * 6 + 8)
    # SURVEY 8(d)'s model of the same two kernels (what the reference's algorithm would move)
    model_scan = model["windows"] * 32.0 + 40.0
    model_ext = 32.0 * model["queries"] + model["compares"] * (16 + 32.0) + 17

    try:
        rnd = json.load(open(os.path.join(ROOT, "profiles", "hbm_random.json")))
    except Exception:
        rnd = None

    def roof(name, kernel_key, ms_, need, model_b, bound_note):
        ach = need * n_reads / (ms_ / 1e3) / 1e9
        tr = measured_traffic(kernel_key)
        tr_gbs = None if tr is None else tr / (ms_ / 1e3) / 1e9
        random_access = None
        if tr_gbs is not None and rnd is not None and large:
            # on an index far larger than L2 the kernels' DRAM traffic is random reads, each of which brings its
            # 128-byte line from HBM: the rate against what a plain random-read kernel makes this GPU's HBM deliver
            # (profiles/hbm_random.json: 36 G random reads/s, whatever their size up to the line)
            ceil = rnd["dram_gbs_when_every_access_fetches_its_128B_line"]
            random_access = {"dram_gbs": tr_gbs, "ceiling_dram_gbs": ceil, "frac_of_ceiling": tr_gbs / ceil,
                             "ceiling_source": rnd["source"]}
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": tr, "traffic_achieved_gbs": tr_gbs,
                "traffic_frac_of_peak": None if tr_gbs is None else tr_gbs / peak, "random_access": random_access,
                "necessary_bytes_per_read": need, "necessary_bytes_per_launch": need * n_reads, "kernel_ms": ms_,
                "model_bytes_per_read": model_b, "model_frac": model_b * n_reads / (ms_ / 1e3) / 1e9 / peak,
                "binding": bound_note, "peak_source": peak_src}

    tier = "L2/HBM (minimizer-addressed filter)" if large else "shared memory (filter copied into every CTA)"
    k_scan = roof("k_scan_tiles_ws (pack + seed scan / index probe)", "k_scan_tiles", scan_ms, scan_bytes, model_scan,
                  "filter tier: %s. With the filter in shared memory the kernel is bound by the integer ALU pipe (80 %% busy, "
                  "profiles/r01_ncu_k_scan_tiles_v11_5Mbp.txt; 70 %% at 100 Mbp with the filter in L2); on an index far larger "
                  "than L2 by the rate at which HBM serves random reads: about 14 DRAM transactions per read "
                  "(profiles/r02_scan_experiments.txt, profiles/r02b_ncu_k_scan_tiles_3Gbp.txt)" % tier)
    k_ext = roof("k_extend_greedy (overlap extension)" if not EXHAUSTIVE else "k_extend_exhaustive_nodes (overlap extension)",
                 "k_extend_greedy" if not EXHAUSTIVE else "k_extend_exhaustive", ext_ms, ext_bytes, model_ext,
                 "dependent round trips (entry side, then unitig words) at 20 warps per SM and 15 of 32 threads per "
                 "instruction (profiles/r02_ncu_k_extend_greedy_*.txt): latency and issue bound, not bandwidth bound")
    if EXHAUSTIVE:
        # the oracle's per-read queries / comparisons count what the REFERENCE does in this mode -- getEnd for every
        # window of every read (alignerExhaustive.cpp:41) -- most of which the hit mask of the seed scan makes
        # unnecessary here: no figure for this design's necessary bytes, only the measured traffic
        k_ext["frac"] = k_ext["achieved"] = k_ext["necessary_bytes_per_read"] = k_ext["necessary_bytes_per_launch"] = None
    dominant = k_scan if (scan_ms >= ext_ms or EXHAUSTIVE) else k_ext
    roofline = dict(dominant)
    roofline["note"] = ("frac = necessary bytes of THIS design / kernel time / measured HBM peak; model_frac = the same with "
                        "SURVEY 8(d)'s bytes (one 32-byte index sector per window the reference examines), which the "
                        "filter makes unnecessary to move; traffic = DRAM bytes ncu measured for this kernel and graph size, "
                        "traffic_frac_of_peak = that traffic / kernel time / measured HBM peak")
    roofline["kernels"] = {"scan": k_scan, "extend": k_ext}
    roofline["reference_model_per_read"] = model
    roofline["reference_model_source"] = model_src
    cfg = workload_config(args)
    line = {
        "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms[2] / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": cfg,
        "index": {"unitigs": stats["unitigs"], "border_keys": stats["keys"], "bytes": stats["device_bytes"],
                  "build_s": index_build_s, "repeat_copies": int(graph.n_repeat_copies)},
        "e2e_batch_reads": B, "rank0_numa_binding": numa,
        "bases_per_s": value * READ_LEN,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": e2e_forms[e2e_best]["h2d_bytes_per_step"],
                "d2h_bytes_per_step": e2e_forms[e2e_best]["d2h_bytes_per_step"], "ms_per_step": 1e3 * e2e_s_best / steps,
                "upload": e2e_best, "forms": e2e_forms},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernel_ms_per_pass": {"pack_scan": scan_ms, "extend": ext_ms, "whole": dev_ms[2] / per_pass},
        "timed_region_s": {"resident_wall": t_res1 - t_res0, "resident_device": dev_ms[2] / 1e3},
        "packed_input": {"value": world * n_reads * per_pass / (dev_pk[2] / 1e3), "unit": "reads/s",
                         "kernel_ms_per_pass": {"scan": dev_pk[0] / per_pass, "extend": dev_pk[1] / per_pass},
                         "what": "same resident leg with the reads in HBM as the 2-bit stream of a packed upload"},
        "outcome": {"scope": "all ranks, per pass", "reads": counters["reads"] // per_pass, "aligned": counters["aligned"] // per_pass,
                    "not_aligned": counters["not_aligned"] // per_pass, "no_overlap": counters["no_overlap"] // per_pass,
                    "filter_passes_per_read": verified_per_read, "mapped_e2e_last_pass": mapped_e2e},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if world == 1 and not args.no_cli:
        al.close()  # the command line creates its own context
        line["cli"] = cli_leg(graph, args.cli_reads)
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
