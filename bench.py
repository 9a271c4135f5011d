#!/usr/bin/env python
"""bench.py - mapped reads/s of the PEMapper hot path (seed lookup -> Smith-Waterman -> pileup) on N B200s.

Workload (default, BASELINE.json configs[2] = the north-star target): synthetic human-sized genome, 24 contigs with
sizes proportional to chr1-22, X, Y totalling 3.1e9 bp, paired-end 150 bp reads with 1 % substitutions and 0.1 % 1-3 bp
insertions / deletions, insert U[250,450].  One "step" = one pass of the hot path over 10 M pairs (20 M read-mates) per
GPU.  `--config cfg2` runs BASELINE.json configs[1] (64 Mb single contig) instead.  With N > 1 ranks the reads are
sharded weakly (every rank maps its own 10 M pairs against its replica of the index) and the step ends with the sum of
the per-GPU pileup counter arrays onto rank 0 (NCCL over NVLink, chromosome by chromosome).

  value  reads/s with the reads already in HBM (pemap_map_batch_device), CUDA events on the library's stream
  e2e    reads/s through the C-ABI with pinned HOST buffers: H2D of the reads, D2H of m1/m2/type AND one
         pemap_finish_stream (bounded compaction + pinned D2H of every pileup record of the step) inside the timed
         region, counters reset before every step
  --impl reference : the reference's own CPU implementation (oracle/_ref/libpemapper_ref.so = unmodified
         pemapper.c built in-process; falls back to the oracle port) on all of the box's host threads, each step a
         bounded sample (threads x 20,000 pairs, so that every worker thread holds a full batch).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GENOME_LEN = int(os.environ.get("PEMAP_BENCH_GENOME", 64_000_000))  # cfg2: 64 Mb
PAIRS = 10_000_000
READ_LEN = 150
STRIDE = 160
SUB, INS, DEL = 0.01, 0.001, 0.001
INSERT = (250, 450)
MIN_ALIGN, MAX_DIST, MIN_DIST = 0.85, 500, 0
ALG_BYTES_PER_READ_150 = 7840          # 2 strands x 10 segments x 49 k-mers x 8 B (pos_index words)
INT_OPS_PER_CELL = 10
REF_BATCH = 20_000                     # reads_per_thread of the reference (pemapper.c:158)
CHR_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]

CONFIGS = {
    "cfg3": {"genome_seed": 30, "read_seed": 31, "total": float(os.environ.get("PEMAP_CFG3_BASES", 3.1e9)),
             "text": "cfg3: %d contigs (sizes ~ human chr1-22,X,Y) totalling %.2e bp"},
    "cfg2": {"genome_seed": 20, "read_seed": 21, "total": float(GENOME_LEN), "text": "cfg2: %d contig of %.2e bp"},
    "cfg5": {"genome_seed": 50, "read_seed": 51, "total": 64e6,
             "text": "cfg5: %d contigs totalling %.2e bp, half of it copies of 200 2-kb units at 0-2 %% divergence"},
}


def env_int(k, d):
    return int(os.environ.get(k, d))


# ------------------------------------------------------------------------------------------ synthetic data (torch)

def make_genome(seed, n):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]


def config_genome(name, device):
    """-> (list of host contigs (uint8 arrays), the concatenated genome as a torch tensor on `device`).
    cfg2 uses numpy's PCG64 stream (the tests' genome); cfg3's 3.1e9 bases come from torch's generator on `device`."""
    import torch
    cfg = CONFIGS[name]
    if name == "cfg2":
        g = make_genome(cfg["genome_seed"], int(cfg["total"]))
        return [g], torch.from_numpy(g).to(device)
    if name == "cfg5":
        from pecaller_b200 import synth
        contigs = synth.repeat_genome(cfg["genome_seed"], [int(cfg["total"]) // 16] * 16, unit_len=2000, n_units=200, frac=0.5,
                                      max_div=0.02)
        return contigs, torch.from_numpy(np.concatenate(contigs)).to(device)
    lens = [int(cfg["total"] * m / sum(CHR_MB)) for m in CHR_MB]
    G = sum(lens)
    gen = torch.Generator(device=device)
    gen.manual_seed(cfg["genome_seed"])
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    gt = torch.empty(G, dtype=torch.uint8, device=device)
    step = 1 << 28
    for lo in range(0, G, step):
        m = min(step, G - lo)
        gt[lo:lo + m] = acgt[torch.randint(0, 4, (m,), generator=gen, device=device)]
    host = gt.cpu().numpy()
    contigs, at = [], 0
    for L in lens:
        contigs.append(host[at:at + L])
        at += L
    return contigs, gt


def workload_text(name, contigs, pairs):
    cfg = CONFIGS[name]
    return (cfg["text"] % (len(contigs), float(sum(int(c.shape[0]) for c in contigs))) +
            ", %d paired-end 150 bp reads per GPU per step, 1%% subs + 0.1%% ins + 0.1%% del, insert U[250,450]" % pairs)


def host_index(contigs, params, device_ok):
    """The index as pemapper's main() holds it (pos_index[2^32+1], mers, contig_starts), for the CPU arm.  Built by the
    device index builder when a GPU is there (bit-equal to index_genome_whole's files: tests/test_gpu_parity.py),
    read back through a handle that holds nothing but the index (PEMAP_INDEX_ONLY=1); else by the oracle's indexer."""
    lens = np.array([int(c.shape[0]) for c in contigs], dtype=np.int64)
    cstarts = np.concatenate([[0], np.cumsum(lens - 15)]).astype(np.uint32)
    if device_ok:
        import pecaller_b200 as pb
        os.environ["PEMAP_INDEX_ONLY"] = "1"
        try:
            m = pb.PEMapper.from_genome(contigs, params)
        finally:
            del os.environ["PEMAP_INDEX_ONLY"]
        mers = np.concatenate([m.read_mers(), np.zeros(4, np.uint32)])
        total = (1 << 32) + 1
        pos_index = np.empty(total, dtype=np.uint32)
        step = 1 << 30
        for first in range(0, total, step):
            k = min(step, total - first)
            pos_index[first:first + k] = m.read_pos_index(first, k)
        m.close()
        return pos_index, mers, cstarts
    import oracle_lib as ol
    o = ol.Oracle(contigs, ol.default_params())
    out = (o.dense_pos_index(), np.concatenate([o.mers(), np.zeros(4, np.uint32)]), o.contig_starts())
    o.close()
    return out


def torch_reads(genome_t, n_pairs, seed, device, chunk=1_000_000):
    """Paired reads with the config-2 error model, generated with torch ops on `device` (same model as
    pecaller_b200.synth.simulate_reads; used here because 20 M reads take minutes in numpy)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    comp = torch.full((256,), ord("N"), dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    code = torch.zeros(256, dtype=torch.long, device=device)
    code[ord("C")], code[ord("G")], code[ord("T")] = 1, 2, 3
    G = genome_t.shape[0]
    out1 = torch.zeros((n_pairs, STRIDE), dtype=torch.uint8, device=device)
    out2 = torch.zeros((n_pairs, STRIDE), dtype=torch.uint8, device=device)

    def extract(start, m):
        step = torch.ones((m, READ_LEN), dtype=torch.int64, device=device)
        u = torch.rand((m, READ_LEN), generator=g, device=device)
        dele = u < DEL
        step += dele * torch.randint(1, 4, (m, READ_LEN), generator=g, device=device)
        ins0 = (u > 1.0 - INS)
        ins0[:, READ_LEN - 3:] = False
        k = torch.randint(1, 4, (m, READ_LEN), generator=g, device=device) * ins0
        inserted = ins0.clone()
        inserted[:, 1:] |= k[:, :-1] >= 2
        inserted[:, 2:] |= k[:, :-2] >= 3
        inserted[:, 0] = False
        step[inserted] = 0
        step[:, 0] = 0
        idx = (start[:, None] + torch.cumsum(step, dim=1)).clamp_(max=G - 1)
        rows = genome_t[idx]
        rnd = acgt[torch.randint(0, 4, (m, READ_LEN), generator=g, device=device)]
        rows = torch.where(inserted, rnd, rows)
        sub = torch.rand((m, READ_LEN), generator=g, device=device) < SUB
        alt = acgt[(code[rows.long()] + torch.randint(1, 4, (m, READ_LEN), generator=g, device=device)) & 3]
        return torch.where(sub, alt, rows)

    for lo in range(0, n_pairs, chunk):
        m = min(chunk, n_pairs - lo)
        frag = torch.randint(INSERT[0], INSERT[1] + 1, (m,), generator=g, device=device)
        start = (torch.rand((m,), generator=g, device=device, dtype=torch.float64) * (G - frag - 32)).long()
        left = extract(start, m)
        right = comp[extract(start + frag - READ_LEN, m).flip(1).long()]
        rev = torch.rand((m,), generator=g, device=device) < 0.5
        out1[lo:lo + m, :READ_LEN] = torch.where(rev[:, None], right, left)
        out2[lo:lo + m, :READ_LEN] = torch.where(rev[:, None], left, right)
    return out1, out2


# ------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arms

class CpuArm:
    """The reference's own CPU implementation of the path on a bounded sample: libpemapper_ref.so (the unmodified
    pemapper.c, `kind` = "reference") when it was built, else the oracle port."""

    def __init__(self, contigs, index_arrays, threads):
        import oracle_lib as ol
        self.threads = threads
        self.kind = "port"
        self.ref = None
        self.oracle = None
        if ol.have_reference_lib() and os.environ.get("PEMAP_BENCH_PORT", "0") != "1":
            try:
                pos_index, mers, cstarts = index_arrays()
                cat = contigs[0] if len(contigs) == 1 else np.concatenate(contigs)
                self.ref = ol.ReferenceLib(None, min_align=MIN_ALIGN, paired=True, min_dist=MIN_DIST, max_dist=MAX_DIST,
                                           arrays=(cat, cstarts, pos_index, mers, len(contigs)))
                self.kind = "reference"
            except Exception as e:  # e.g. not enough host RAM for the 32 B x genome BASE_NODE array
                sys.stderr.write("reference library unusable (%s); timing the oracle port\n" % e)
        if self.ref is None:
            self.oracle = ol.Oracle(contigs, ol.default_params(min_align=MIN_ALIGN, pair_flag=1, min_dist=MIN_DIST, max_dist=MAX_DIST))
        self.live = 0

    def run(self, r1, r2, steps, warmup):
        rates = []
        for it in range(warmup + steps):
            t = time.perf_counter()
            if self.ref is not None:
                self.ref.map(r1, r2, nthreads=self.threads)
                self.live = self.ref.live_threads()
            else:
                self.oracle.map_batch(r1, r2, nthreads=self.threads)
                self.live = min(self.threads, (r1.shape[0] + REF_BATCH - 1) // REF_BATCH)
            dt = time.perf_counter() - t
            if it >= warmup:
                rates.append(2 * r1.shape[0] / dt)
        return rates


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def alu_peak():
    """Measured integer / fp64 issue rates (tools/alu_peak.cu on a B200), committed under profiles/."""
    p = os.path.join(ROOT, "profiles", "alu_peak.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


# ------------------------------------------------------------------------------------------ main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("PEMAP_BENCH_CONFIG", "cfg3"), choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=env_int("PEMAP_BENCH_PAIRS", PAIRS), help="pairs per rank per step")
    ap.add_argument("--cpu-sample-pairs", type=int, default=env_int("PEMAP_BENCH_CPU_PAIRS", 0),
                    help="pairs per CPU step (default: host threads x 20,000 = one full batch per worker thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reduce-every", type=int, default=env_int("PEMAP_BENCH_REDUCE_EVERY", 4),
                    help="N > 1: steps between two sums of the per-GPU pileup counters in the device-resident leg.  The "
                         "target run (30x = 310 M pairs) is 31 steps of 10 M pairs, i.e. ~4 per GPU at N = 8 and more at "
                         "smaller N, with ONE sum at its end; the end-to-end leg sums (and hands over the pileup) every step")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    threads = os.cpu_count() or 1
    cpu_pairs = a.cpu_sample_pairs or threads * REF_BATCH
    cfg = CONFIGS[a.config]

    import torch

    if a.impl == "reference":
        if rank != 0:
            return 0
        import pecaller_b200 as pb
        have_gpu = torch.cuda.is_available()
        dev = torch.device("cuda", local) if have_gpu else torch.device("cpu")
        contigs, gt = config_genome(a.config, dev)
        r1, r2 = torch_reads(gt, cpu_pairs, cfg["read_seed"], dev, chunk=250_000)
        r1, r2 = r1[:, :READ_LEN].cpu().numpy(), r2[:, :READ_LEN].cpu().numpy()
        del gt
        if have_gpu:
            torch.cuda.empty_cache()
        params = pb.default_params(min_align=MIN_ALIGN, pair_flag=1, min_dist=MIN_DIST, max_dist=MAX_DIST) if have_gpu else None
        t0 = time.time()
        arm = CpuArm(contigs, lambda: host_index(contigs, params, have_gpu), threads)
        t_setup = time.time() - t0
        rates = arm.run(r1, r2, a.steps, a.warmup)
        v = float(np.mean(rates))
        sample = "%d pairs (%d read-mates) per step = one 20,000-read batch for each of %d worker threads (%d live)" % (
            cpu_pairs, 2 * cpu_pairs, threads, arm.live)
        line = {"impl": "reference", "metric": "mapped reads/sec", "value": v, "unit": "reads/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * 2 * cpu_pairs / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(a.config, contigs, a.pairs),
                           "sample": "each step maps the first %d pairs of that workload on %d host threads" % (cpu_pairs, threads),
                           "setup_s": round(t_setup, 1)},
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "threads_live": arm.live, "kind": arm.kind,
                                 "sample": sample},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    import torch.distributed as dist
    import pecaller_b200 as pb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    params = pb.default_params(min_align=MIN_ALIGN, pair_flag=1, min_dist=MIN_DIST, max_dist=MAX_DIST)
    t0 = time.time()
    contigs, gt = config_genome(a.config, dev)
    t_genome = time.time() - t0
    want_cpu = (not a.no_cpu_baseline) and world == 1
    idx_host = None
    if want_cpu:  # the CPU arm needs the index on the host: fetch it before the mapper fills the HBM
        idx_host = host_index(contigs, params, True)
    n = a.pairs
    t0 = time.time()
    d_r1, d_r2 = torch_reads(gt, n, cfg["read_seed"] + rank, dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    del gt
    torch.cuda.empty_cache()
    t0 = time.time()
    mapper = pb.PEMapper.from_genome(contigs, params, device=local)
    t_index = time.time() - t0
    d_len = torch.full((n,), READ_LEN, dtype=torch.int32, device=dev)
    d_m1 = torch.zeros(n, dtype=torch.int32, device=dev)
    d_m2 = torch.zeros(n, dtype=torch.int32, device=dev)
    d_ty = torch.zeros(n, dtype=torch.int32, device=dev)
    # pinned host copies for the end-to-end leg
    h_r1 = torch.empty((n, STRIDE), dtype=torch.uint8, pin_memory=True).copy_(d_r1)
    h_r2 = torch.empty((n, STRIDE), dtype=torch.uint8, pin_memory=True).copy_(d_r2)
    h_len = torch.full((n,), READ_LEN, dtype=torch.int32).pin_memory()
    h_m1 = torch.zeros(n, dtype=torch.int32).pin_memory()
    h_m2 = torch.zeros(n, dtype=torch.int32).pin_memory()
    h_ty = torch.zeros(n, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()
    free_b, total_b = torch.cuda.mem_get_info()

    lib_stream = torch.cuda.ExternalStream(mapper.stream_ptr(), device=dev)
    from pecaller_b200 import sharding
    reducer = sharding.SliceReducer(mapper) if world > 1 else None
    site_range = [None]
    red = {"s": 0.0, "n": 0}   # wall time and count of the slice-wise counter sums inside the timed device leg

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i=0, last=True):
        mapper.map_device(n, d_r1.data_ptr(), d_len.data_ptr(), d_r2.data_ptr(), d_len.data_ptr(), STRIDE, READ_LEN,
                          d_m1.data_ptr(), d_m2.data_ptr(), d_ty.data_ptr())
        if world > 1 and (last or (i + 1) % a.reduce_every == 0):
            # slice-wise sum over NVLink peer memory: rank r ends up with the final counters of its 1/N of the genome
            t_r = time.perf_counter()
            site_range[0] = reducer.reduce_scatter()
            red["s"] += time.perf_counter() - t_r
            red["n"] += 1
            mapper.reset_counts()

    fin = {"records": 0}

    def step_host():
        mapper.reset_counts()
        mapper._ck(mapper._L.pemap_map_batch_rows(mapper._h, n, h_r1.data_ptr(), h_len.data_ptr(), h_r2.data_ptr(),
                                                  h_len.data_ptr(), STRIDE, h_m1.data_ptr(), h_m2.data_ptr(), h_ty.data_ptr()))
        if world > 1:
            site_range[0] = reducer.reduce_scatter()
        # the writer's input: every covered site of the step's pileup crosses to pinned host memory (N > 1: every rank
        # compacts and copies out its own slice of the genome, in parallel)
        fin["records"] = mapper.finish_stream(None, site_range=site_range[0])

    # ---- device-resident leg
    for _ in range(a.warmup):
        mapper.reset_counts()
        step_device()
    mapper.reset_counts()
    mapper.reset_stats()
    red["s"], red["n"] = 0.0, 0
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(lib_stream)
    t0 = time.perf_counter()
    for i in range(a.steps):
        step_device(i, i == a.steps - 1)
    ev1.record(lib_stream)
    barrier()
    wall_dev = time.perf_counter() - t0
    ms_dev = ev0.elapsed_time(ev1)
    if world > 1:  # the barriers around the slice-wise sum are host-side: use the wall clock there
        ms_dev = max(ms_dev, 1000.0 * wall_dev)
    stats = mapper.stats()
    clocks = sampler.stop()
    mapped = int((d_m1 != 0).sum().item() + (d_m2 != 0).sum().item())
    types = torch.bincount(d_ty.long(), minlength=9).tolist()

    # ---- end-to-end leg (host buffers, copies and the pileup hand-over inside the timed region)
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_host()
    barrier()
    wall_e2e = time.perf_counter() - t0
    same = bool((h_m1.to(dev) == d_m1).all().item() and (h_m2.to(dev) == d_m2).all().item())

    # ---- the same end-to-end step with 2-bit packed rows on the host (what the C host's decoder threads can emit):
    # 64 bytes per read instead of 160 cross PCIe; packing is done once, outside the timed region
    from pecaller_b200 import mapper as _m
    pstride = int(mapper._L.pemap_packed_stride(READ_LEN))
    hp = [torch.empty((n, pstride), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for src, dst in ((h_r1, hp[0]), (h_r2, hp[1])):
        for lo in range(0, n, 1_000_000):
            hi = min(n, lo + 1_000_000)
            dst[lo:hi] = torch.from_numpy(_m.pack_reads(src[lo:hi, :READ_LEN].numpy(), None, READ_LEN)[0])

    def step_packed():
        mapper.reset_counts()
        mapper._ck(mapper._L.pemap_map_batch_packed(mapper._h, n, hp[0].data_ptr(), h_len.data_ptr(), hp[1].data_ptr(),
                                                    h_len.data_ptr(), READ_LEN, h_m1.data_ptr(), h_m2.data_ptr(), h_ty.data_ptr()))
        if world > 1:
            site_range[0] = reducer.reduce_scatter()
        fin["records"] = mapper.finish_stream(None, site_range=site_range[0])
    step_packed()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_packed()
    barrier()
    wall_packed = time.perf_counter() - t0
    same_packed = bool((h_m1.to(dev) == d_m1).all().item() and (h_m2.to(dev) == d_m2).all().item())
    del hp

    t = torch.tensor([ms_dev, 1000.0 * wall_e2e, 1000.0 * wall_packed], dtype=torch.float64, device=dev)
    nrec = torch.tensor([fin["records"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nrec, op=dist.ReduceOp.SUM)
    ms_dev, ms_e2e, ms_packed = t.tolist()
    fin["records"] = int(nrec.item())
    reads_per_step = 2 * n * world
    value = reads_per_step * a.steps / (ms_dev / 1000.0)
    e2e = reads_per_step * a.steps / (ms_e2e / 1000.0)

    if rank == 0:
        hbm_peak, peak_src, _ = peaks()
        per_step = {k: stats[k] / a.steps for k in ("lookups", "mer_positions", "candidates", "sw_cells", "tb_cells",
                                                     "tb_cells_int", "sw_cells_certified")}
        # cells the DP kernel really computed: candidates settled by the ungapped-diagonal certificate are not counted
        per_step["sw_cells_dp"] = per_step["sw_cells"] - per_step["sw_cells_certified"]
        seed_bytes = per_step["lookups"] * 8 + per_step["mer_positions"] * 4 + 2 * n * ((READ_LEN + 3) // 4)
        seed_s = stats["ms_seed"] / a.steps / 1000.0
        sw_s = stats["ms_sw"] / a.steps / 1000.0
        tb_s = stats["ms_traceback"] / a.steps / 1000.0
        tbi_s = max(stats["ms_tb_int"] / a.steps / 1000.0, 1e-9)
        tbf_s = max(stats["ms_tb_fp64"] / a.steps / 1000.0, 1e-9)
        seed_gbs = seed_bytes / seed_s / 1e9
        sw_gcups = per_step["sw_cells_dp"] / sw_s / 1e9
        ap_ = alu_peak() or {}
        pk16, pk32, pk64 = ap_.get("sw_s16x2_gcups_peak"), ap_.get("sw_s32_gcups_peak"), ap_.get("sw_fp64_gcups_peak")
        src = ("profiles/alu_peak.json (issue rates measured with tools/alu_peak.cu on a B200; ceiling = the inner loop's issue "
               "budget: per pair of packed cells 2 half-rate DPX + 8 single-rate ops)")

        def alu_roof(kernel, cells, secs, peak):
            g = cells / secs / 1e9
            return {"kernel": kernel, "bound": "alu", "achieved": g, "peak": peak, "unit": "GCUPS",
                    "frac": (g / peak) if peak else None, "traffic": None, "ms_per_step": 1000 * secs,
                    "peak_source": src if peak else "not measured"}
        roof_seed = {"kernel": "k_seed_rbi", "bound": "hbm", "achieved": seed_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": seed_gbs / hbm_peak, "traffic": None, "peak_source": peak_src, "ms_per_step": 1000 * seed_s,
                     "glookups_per_s": per_step["lookups"] / seed_s / 1e9,
                     "note": "achieved = SURVEY 8d's ALGORITHMIC bytes (8 B per pos_index lookup of the reference's 49 per "
                             "segment + 4 B per gathered position + packed read) over the stage time.  The kernel does not "
                             "do those lookups one by one (isolated 8-byte gathers top out at 41.7 G/s = 334 GB/s of such "
                             "bytes on this part, profiles/gather_probe_r01.json): it streams the four ~1 KB buckets of the "
                             "rotated bucket index that hold a segment's 49 k-mers, so its DRAM traffic (`traffic`, ncu) "
                             "is several times the algorithmic bytes by design and is what runs near the copy bandwidth."}
        roof_sw = alu_roof("k_sw_i16 (s16x2 DPX scoring of the candidates k_diag_certify could not settle)",
                           per_step["sw_cells_dp"], sw_s, pk16)
        roof_sw["note"] = ("cells computed by the DP only; %.1f %% of the candidates' cells were settled by the "
                           "ungapped-diagonal certificate (k_diag_certify, inside this stage's time)"
                           % (100.0 * per_step["sw_cells_certified"] / max(per_step["sw_cells"], 1)))
        roof_tbi = alu_roof("k_trace_dp16 (s16x2 integer traceback of gapped winners: k_trace_dp16 + k_trace_walk16)",
                            per_step["tb_cells_int"], tbi_s, pk16)
        roof_tbf = alu_roof("k_sw_fp64 (exact traceback after a rational tie)", per_step["tb_cells"], tbf_s, pk64)
        dominant = max((roof_seed, roof_sw, roof_tbi, roof_tbf), key=lambda r: r["ms_per_step"])
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            trj = json.load(open(tr)).get(a.config, {})
            for rf in (roof_seed, roof_sw, roof_tbi, roof_tbf):
                rf["traffic"] = trj.get(rf["kernel"].split(" ")[0])
        line = {"metric": "mapped reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "s16x2 integer DP in units of 1/36 (fp64 only for rational ties)",
                "data": "synthetic",
                "config": {"workload": workload_text(a.config, contigs, n),
                           "l2": "inputs (%.1f GB of reads per step) and the index exceed the 126 MB L2" % (2 * n * STRIDE / 1e9),
                           "parallelism": ("reads sharded over %d GPUs, index replicated, pileup counters summed slice-wise over NVLink peer "
                                           "memory every %d steps (and at the end of the timed region), every GPU compacts its own "
                                           "1/%d of the genome" % (world, a.reduce_every, world)) if world > 1 else "1 GPU",
                           "counter_sum_ms": round(1000.0 * red["s"] / max(red["n"], 1), 1) if world > 1 else None,
                           "counter_sums_in_timed_region": red["n"],
                           "index_build_s": round(t_index, 2), "datagen_s": round(t_gen + t_genome, 2),
                           "hbm_used_gb": round((total_b - free_b) / 1e9, 1)},
                "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": 2 * n * STRIDE + 2 * n * 4,
                        "d2h_bytes_per_step": 3 * n * 4 + 16 * fin["records"], "ms_per_step": ms_e2e / a.steps,
                        "matches_device_leg": same, "pileup_records_per_step": fin["records"],
                        "includes": "pemap_reset_counts + pemap_map_batch_rows (pinned rows)" +
                                    (" + slice-wise counter sum" if world > 1 else "") + " + pemap_finish_stream per step"},
                "e2e_packed_reads": {"value": reads_per_step * a.steps / (ms_packed / 1000.0), "unit": "reads/s",
                                     "h2d_bytes_per_step": 2 * n * pstride + 2 * n * 4,
                                     "d2h_bytes_per_step": 3 * n * 4 + 16 * fin["records"], "ms_per_step": ms_packed / a.steps,
                                     "matches_device_leg": same_packed,
                                     "note": "pemap_map_batch_packed: 2-bit codes + N mask, %d bytes per read" % pstride},
                "gpu_launches": int(stats["launches"]),
                "clocks": clocks,
                "roofline": dominant, "roofline_seed": roof_seed, "roofline_sw": roof_sw,
                "roofline_traceback_int": roof_tbi, "roofline_traceback_fp64": roof_tbf,
                "sw_gcups": sw_gcups, "seed_gather_gbs": seed_gbs,
                "stage_ms_per_step": {"seed": 1000 * seed_s, "sw": 1000 * sw_s, "select": stats["ms_select"] / a.steps,
                                      "traceback": 1000 * tb_s, "traceback_diag": stats["ms_tb_diag"] / a.steps,
                                      "traceback_int": 1000 * tbi_s, "traceback_fp64": 1000 * tbf_s,
                                      "sum_of_chunks": stats["ms_total"] / a.steps},
                "mapped_reads_per_step": mapped, "mapping_types": types,
                "tracebacks_per_step": {"pure_diagonal": stats["diag_traced"] / a.steps,
                                        "fp64_after_tie": stats["exact_traced"] / a.steps,
                                        "replayed_read_mates": stats["replayed"] / a.steps}}
    r1 = h_r1[:cpu_pairs, :READ_LEN].numpy().copy() if want_cpu else None
    r2 = h_r2[:cpu_pairs, :READ_LEN].numpy().copy() if want_cpu else None
    mapper.close()
    del h_r1, h_r2, d_r1, d_r2
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        if want_cpu:  # after the GPU legs and with their buffers released: the reference wants 32 B per genome base
            arm = CpuArm(contigs, lambda: idx_host, threads)
            rates = arm.run(r1, r2, 1, 0)
            line["cpu_baseline"] = {"value": rates[0], "unit": "reads/s", "cores": threads, "threads_live": arm.live,
                                    "kind": arm.kind,
                                    "sample": "first %d pairs (%d read-mates) of the same workload = one 20,000-read batch "
                                              "for each of %d worker threads, 1 pass" % (cpu_pairs, 2 * cpu_pairs, threads)}
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
