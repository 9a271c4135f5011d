"""CPU: the C host's own logic (pecaller_b200/host/pemapper_gpu.c), exercised without a GPU by host_selftest.c:
the FASTQ reader / decoder against the reference's record scan (pemapper.c:713-739, 663) on plain and gzip input with
'@' quality lines, the short-read rule and trimming; the parallel gzip writer's members must inflate into the record
stream byte for byte."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_reader_and_gzip_writer(tmp_path):
    from pecaller_b200 import build as b
    b.build()                                              # the host links the C-ABI library (no GPU call is made)
    host = os.path.join(ROOT, "pecaller_b200", "host")
    exe = str(tmp_path / "host_selftest")
    subprocess.run(["gcc", "-O2", "-Wno-unused-result", "-Wno-unused-function", "-Wno-format-truncation",
                    "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(host, "host_selftest.c"),
                    "-L", os.path.join(ROOT, "pecaller_b200"), "-lpemap", "-lz", "-lpthread",
                    "-Wl,-rpath," + os.path.join(ROOT, "pecaller_b200")], check=True)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok ") == 3, r.stdout
