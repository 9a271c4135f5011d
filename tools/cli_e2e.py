#!/usr/bin/env python
"""FASTQ-to-files rate of the C host (pecaller_b200/host/pemapper_gpu) on one B200: cfg2's genome and error model,
N pairs written as plain and as gzip FASTQ, mapped with the reference's command line.  Reports the CLI's own stage
timers (PEMAP_TIMING=1: decode, GPU map, gz writer), the wall clock and the reads/s from FASTQ to closed output files.
    python tools/cli_e2e.py [pairs] > profiles/cli_e2e_r02.json"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def write_fastq_fixed(path, reads):
    """fixed-width records (12-byte header line) so that the file is one numpy matrix"""
    n, L = reads.shape
    rec = np.empty((n, 12 + L + 3 + L + 1), dtype=np.uint8)
    ids = np.arange(n)
    rec[:, 0] = ord("@")
    rec[:, 1] = ord("r")
    for k in range(9):
        rec[:, 10 - k] = ord("0") + (ids // 10 ** k) % 10
    rec[:, 11] = 10
    rec[:, 12:12 + L] = reads
    rec[:, 12 + L:12 + L + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 15 + L:15 + 2 * L] = ord("I")
    rec[:, 15 + 2 * L] = 10
    rec.tofile(path)


def run_cli(exe, out, sdx, f1, f2, n, env):
    cmd = [exe, out, sdx, "p", f1, f2, "500", "0", "n", "0.85", str(os.cpu_count() or 8), str(n)]
    t0 = time.time()
    r = subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    timing = {}
    for ln in r.stderr.splitlines():
        if ln.startswith("{"):
            timing = json.loads(ln)
    return wall, timing


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    n_gz = min(n, int(os.environ.get("PEMAP_CLI_GZ_PAIRS", 2_000_000)))
    dev = torch.device("cuda", 0)
    contigs, gt = bench.config_genome("cfg2", dev)
    d_r1, d_r2 = bench.torch_reads(gt, n, 21, dev)
    r1 = d_r1[:, :bench.READ_LEN].cpu().numpy()
    r2 = d_r2[:, :bench.READ_LEN].cpu().numpy()
    del gt, d_r1, d_r2
    torch.cuda.empty_cache()
    tmp = tempfile.mkdtemp(prefix="pemap_cli_", dir=os.environ.get("TMPDIR", "/tmp"))
    try:
        g = contigs[0]
        with open(os.path.join(tmp, "g.sdx"), "w") as f:
            f.write("1\n%d\tchr1\n16\n" % (g.shape[0] - 15))
        with gzip.open(os.path.join(tmp, "g.seq"), "wb", compresslevel=1) as f:
            f.write(g.tobytes())
        write_fastq_fixed(os.path.join(tmp, "r_1.fq"), r1)
        write_fastq_fixed(os.path.join(tmp, "r_2.fq"), r2)
        for k in (1, 2):
            with open(os.path.join(tmp, "r_%d.fq" % k), "rb") as fi, gzip.open(os.path.join(tmp, "z_%d.fq.gz" % k), "wb", compresslevel=4) as fo:
                shutil.copyfileobj(_Limited(fi, n_gz * (12 + 2 * bench.READ_LEN + 4)), fo, 1 << 24)
        exe = os.path.join(ROOT, "pecaller_b200", "host", "pemapper_gpu")
        env = dict(os.environ, PEMAP_DEVICE_INDEX="1", PEMAP_TIMING="1")
        out = {"pairs": n, "read_len": bench.READ_LEN, "host_threads": os.cpu_count(), "runs": []}
        for name, f1, f2, m, lvl in (("plain FASTQ, gz level 6", "r_1.fq", "r_2.fq", n, "6"), ("plain FASTQ, gz level 1", "r_1.fq", "r_2.fq", n, "1"),
                                     ("gzip FASTQ, gz level 1", "z_1.fq.gz", "z_2.fq.gz", n_gz, "1")):
            env["PEMAP_GZ_LEVEL"] = lvl
            wall, timing = run_cli(exe, os.path.join(tmp, "out"), os.path.join(tmp, "g.sdx"), os.path.join(tmp, f1), os.path.join(tmp, f2), m, env)
            out["runs"].append({"input": name, "pairs": m, "wall_s_including_index_build_and_cuda_init": round(wall, 2),
                                "pileup_gz_bytes": os.path.getsize(os.path.join(tmp, "out.pileup.gz")), "cli_timers": timing})
        print(json.dumps(out, indent=1))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


class _Limited:
    def __init__(self, f, limit):
        self.f, self.left = f, limit

    def read(self, n=-1):
        if self.left <= 0:
            return b""
        b = self.f.read(min(n, self.left) if n > 0 else self.left)
        self.left -= len(b)
        return b


if __name__ == "__main__":
    main()
