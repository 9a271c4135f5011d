"""Readers for the committed reference outputs under tests/golden (written by tools/make_golden.py)."""
import gzip
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def have(fixture, run=None):
    p = os.path.join(GOLD, fixture)
    return os.path.exists(os.path.join(p, "index.json")) and (run is None or os.path.exists(os.path.join(p, run + ".pileup.json")))


def index_meta(fixture):
    return json.load(open(os.path.join(GOLD, fixture, "index.json")))


def mfile(fixture, run, k):
    p = os.path.join(GOLD, fixture, "%s.mfile%d.gz" % (run, k))
    if not os.path.exists(p):
        return None
    return np.frombuffer(gzip.open(p).read(), dtype=np.uint32)


def pileup_meta(fixture, run):
    return json.load(open(os.path.join(GOLD, fixture, run + ".pileup.json")))


def summary(fixture, run):
    """-> (dict of mapping-type name -> count, header numbers)"""
    txt = open(os.path.join(GOLD, fixture, run + ".summary.txt")).read()
    counts = {}
    for ln in txt.split("Mapping Type\tCount\tFraction")[1].strip().split("\n"):
        name, cnt, _ = ln.split("\t")
        counts[name] = int(cnt)
    head = [ln for ln in txt.split("\n") if ln.startswith("Total Number")][0].split("\t")
    return counts, {"total_reads": int(head[1]), "avg_len": float(head[3]), "avg_depth": float(head[5]),
                    "avg_insert": float(head[7])}


def indel_lines(fixture, run):
    """normalised .indel.txt lines -> list of (contig, pos, ref, tot, ref_reads, dels, no_ins, sorted insertion strings)"""
    txt = gzip.open(os.path.join(GOLD, fixture, run + ".indel.norm.txt.gz"), "rt").read()
    out = []
    for ln in txt.split("\n")[1:]:
        if ln:
            p = ln.split("\t")
            out.append((p[0], int(p[1]), p[2], int(p[3]), int(p[4]), int(p[5]), int(p[6]), tuple(p[7:])))
    return out


SINGLE_NAMES = {2: "Unique Mapping", 7: "Non-Unique Mapping, discarded", 8: "No mapping reaches threshold"}
PAIR_NAMES = {0: "Unique Mate-Paired", 1: "Unique Mate-Paired with slip", 2: "Unique Single End", 3: "Unique Mis-size",
              4: "Non-Unique Mate-Paired", 5: "Non-Unique Mis-size", 6: "Fragment Mismatch", 7: "Non-unique with no map",
              8: "Neither Map"}
