#!/usr/bin/env python
"""bench.py - mapped reads/s of the PEMapper hot path (seed lookup -> Smith-Waterman -> pileup) on N B200s.

Workload (BASELINE.json configs[1]): synthetic 64 Mb single-contig genome, 10 M paired-end 150 bp reads with
1 % substitutions and 0.1 % 1-3 bp insertions / deletions, insert U[250,450].  One "step" = one pass of the
hot path over the 10 M pairs (20 M read-mates).  With N > 1 ranks the reads are sharded weakly (every rank
maps its own 10 M pairs against its replica of the index) and the step ends with the NCCL sum of the per-GPU
pileup counter arrays onto rank 0.

  value  reads/s with the reads already in HBM (pemap_map_batch_device), CUDA events on the library's stream
  e2e    reads/s through pemap_map_batch_rows with pinned HOST buffers: H2D of the reads and D2H of m1/m2/type
         inside the timed region
  --impl reference : the reference's own CPU implementation (oracle/_ref/libpemapper_ref.so = unmodified
         pemapper.c built in-process; falls back to the oracle port) on the box's host cores, bounded sample.
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


def env_int(k, d):
    return int(os.environ.get(k, d))


# ------------------------------------------------------------------------------------------ synthetic data (torch)

def make_genome(seed, n):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]


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

def cpu_reference_run(genome, r1, r2, threads, steps, warmup, want_setup=True):
    """Time the reference's own CPU implementation of the path on a bounded sample. -> (reads/s list, kind)"""
    import oracle_lib as ol
    oracle = ol.Oracle([genome], ol.default_params(min_align=MIN_ALIGN, pair_flag=1, min_dist=MIN_DIST, max_dist=MAX_DIST))
    kind = "port"
    ref = None
    if ol.have_reference_lib() and os.environ.get("PEMAP_BENCH_PORT", "0") != "1":
        try:
            ref = ol.ReferenceLib(oracle, min_align=MIN_ALIGN, paired=True, min_dist=MIN_DIST, max_dist=MAX_DIST)
            kind = "reference"
        except Exception as e:  # e.g. not enough host RAM for the 16 GiB table
            sys.stderr.write("reference library unusable (%s); timing the oracle port\n" % e)
    rates = []
    for it in range(warmup + steps):
        t = time.perf_counter()
        if ref is not None:
            ref.map(r1, r2, nthreads=threads)
        else:
            oracle.map_batch(r1, r2, nthreads=threads)
        dt = time.perf_counter() - t
        if it >= warmup:
            rates.append(2 * r1.shape[0] / dt)
    return rates, kind


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
    ap.add_argument("--pairs", type=int, default=env_int("PEMAP_BENCH_PAIRS", PAIRS), help="pairs per rank per step")
    ap.add_argument("--cpu-sample-pairs", type=int, default=env_int("PEMAP_BENCH_CPU_PAIRS", 100_000))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    threads = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return 0
        genome = make_genome(20, GENOME_LEN)
        import torch
        gt = torch.from_numpy(genome)
        n = a.cpu_sample_pairs
        r1, r2 = torch_reads(gt, n, 21, "cpu", chunk=250_000)
        r1, r2 = r1[:, :READ_LEN].numpy(), r2[:, :READ_LEN].numpy()
        rates, kind = cpu_reference_run(genome, r1, r2, threads, a.steps, a.warmup)
        v = float(np.mean(rates))
        line = {"impl": "reference", "metric": "mapped reads/sec", "value": v, "unit": "reads/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * 2 * n / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "cfg2: %d Mb single-contig genome, %d paired-end 150 bp reads per GPU per step, "
                                       "1%% subs + 0.1%% ins + 0.1%% del, insert U[250,450]" % (GENOME_LEN // 1000000, a.pairs),
                           "sample": "each step maps the first %d pairs of that workload on %d host threads" % (n, threads)},
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": kind,
                                 "sample": "%d pairs (%d read-mates) per step, %d threads" % (n, 2 * n, threads)},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    import pecaller_b200 as pb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    genome = make_genome(20, GENOME_LEN)
    params = pb.default_params(min_align=MIN_ALIGN, pair_flag=1, min_dist=MIN_DIST, max_dist=MAX_DIST)
    t0 = time.time()
    mapper = pb.PEMapper.from_genome([genome], params, device=local)
    t_index = time.time() - t0
    gt = torch.from_numpy(genome).to(dev)
    n = a.pairs
    t0 = time.time()
    d_r1, d_r2 = torch_reads(gt, n, 21 + rank, dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    del gt
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

    lib_stream = torch.cuda.ExternalStream(mapper.stream_ptr(), device=dev)
    from pecaller_b200 import sharding
    counts_t = sharding.counts_tensor(mapper, dev) if world > 1 else None  # torch view of the library's counter array

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        mapper.map_device(n, d_r1.data_ptr(), d_len.data_ptr(), d_r2.data_ptr(), d_len.data_ptr(), STRIDE, READ_LEN,
                          d_m1.data_ptr(), d_m2.data_ptr(), d_ty.data_ptr())
        if world > 1:
            sharding.reduce_counts(counts_t, dst=0)

    def step_host():
        mapper._ck(mapper._L.pemap_map_batch_rows(mapper._h, n, h_r1.data_ptr(), h_len.data_ptr(), h_r2.data_ptr(),
                                                  h_len.data_ptr(), STRIDE, h_m1.data_ptr(), h_m2.data_ptr(), h_ty.data_ptr()))
        if world > 1:
            sharding.reduce_counts(counts_t, dst=0)

    # ---- device-resident leg
    for _ in range(a.warmup):
        mapper.reset_counts()
        step_device()
    mapper.reset_counts()
    mapper.reset_stats()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(lib_stream)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_device()
    ev1.record(lib_stream)
    barrier()
    wall_dev = time.perf_counter() - t0
    ms_dev = ev0.elapsed_time(ev1)
    if world > 1:  # the NCCL reduce runs on torch's stream after the library's blocking call: use the wall clock there
        ms_dev = max(ms_dev, 1000.0 * wall_dev)
    stats = mapper.stats()
    clocks = sampler.stop()
    # size-independent sanity at full size: counter mass == mapped bases (minus N/indel columns handled separately)
    mapped = int((d_m1 != 0).sum().item() + (d_m2 != 0).sum().item())
    types = torch.bincount(d_ty.long(), minlength=9).tolist()

    # ---- end-to-end leg (host buffers, copies inside the timed region)
    mapper.reset_counts()
    step_host()
    mapper.reset_counts()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_host()
    barrier()
    wall_e2e = time.perf_counter() - t0
    same = bool((h_m1.to(dev) == d_m1).all().item() and (h_m2.to(dev) == d_m2).all().item())

    t = torch.tensor([ms_dev, 1000.0 * wall_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
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
        src = "profiles/alu_peak.json (issue rates measured with tools/alu_peak.cu on a B200, SURVEY 8d: 10 ops per cell)"

        def alu_roof(kernel, cells, secs, peak):
            g = cells / secs / 1e9
            return {"kernel": kernel, "bound": "alu", "achieved": g, "peak": peak, "unit": "GCUPS",
                    "frac": (g / peak) if peak else None, "traffic": None, "ms_per_step": 1000 * secs,
                    "peak_source": src if peak else "not measured"}
        roof_seed = {"kernel": "k_seed_chain", "bound": "hbm", "achieved": seed_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": seed_gbs / hbm_peak, "traffic": None, "peak_source": peak_src, "ms_per_step": 1000 * seed_s,
                     "note": "algorithmic bytes (SURVEY 8d: 8 B per pos_index lookup + 4 B per position + packed read) over the "
                             "stage time; random 8-byte gathers from HBM top out at %.1f G lookups/s = %.0f GB/s of algorithmic "
                             "bytes on this part (profiles/gather_probe_r01.json), the L2-resident k-mer filter is how the "
                             "kernel gets past that" % (ap_.get("random_gather_glookups_s", 41.7),
                                                        8 * ap_.get("random_gather_glookups_s", 41.7))}
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
            trj = json.load(open(tr))
            for rf in (roof_seed, roof_sw, roof_tbi, roof_tbf):
                rf["traffic"] = trj.get(rf["kernel"].split(" ")[0])
        line = {"metric": "mapped reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "s16x2 integer DP in units of 1/36 (fp64 only for rational ties)",
                "data": "synthetic",
                "config": {"workload": "cfg2: %d Mb single-contig genome, %d paired-end 150 bp reads per GPU per step, "
                                       "1%% subs + 0.1%% ins + 0.1%% del, insert U[250,450]" % (GENOME_LEN // 1000000, n),
                           "l2": "inputs (%.1f GB of reads per step) and the 16 GiB index exceed the 126 MB L2" % (2 * n * STRIDE / 1e9),
                           "parallelism": "reads sharded over %d GPU(s), index replicated, NCCL sum of pileup counters" % world,
                           "index_build_s": round(t_index, 2), "datagen_s": round(t_gen, 2)},
                "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": 2 * n * STRIDE + 2 * n * 4,
                        "d2h_bytes_per_step": 3 * n * 4, "ms_per_step": ms_e2e / a.steps, "matches_device_leg": same},
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
        if not a.no_cpu_baseline and world == 1:
            ns = a.cpu_sample_pairs
            r1 = h_r1[:ns, :READ_LEN].numpy()
            r2 = h_r2[:ns, :READ_LEN].numpy()
            rates, kind = cpu_reference_run(genome, r1, r2, threads, 1, 0)
            line["cpu_baseline"] = {"value": rates[0], "unit": "reads/s", "cores": threads, "kind": kind,
                                    "sample": "first %d pairs (%d read-mates) of the same workload, %d threads, 1 pass" % (ns, 2 * ns, threads)}
        print(json.dumps(line), flush=True)
    mapper.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
