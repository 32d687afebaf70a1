#!/usr/bin/env python3
"""Splits the packed output of `sagephy <App> -n N ...` (one file per stream + "<file>.idx") into one file per replicate,
named as N separate reference runs with prefixes <prefix>_0, <prefix>_1, ... would have named them — the layout the
reference's partial.py expects (one tree / leaf map per file).

usage: python tools/unpack_batch.py <prefix> [--first-index K] [--out-dir DIR]
"""
import argparse
import glob
import os

SUFFIXES = (".unpruned.tree", ".pruned.tree", ".unpruned.info", ".pruned.info", ".unpruned.guest2host", ".pruned.guest2host",
            ".unpruned.leafmap", ".pruned.leafmap", ".info")


def unpack(prefix: str, first_index: int = None, out_dir: str = None) -> int:
    """Returns the number of files written."""
    written = 0
    for idx_path in sorted(glob.glob(glob.escape(prefix) + "*.idx")):
        packed = idx_path[:-4]
        suffix = packed[len(prefix):]
        if not os.path.exists(packed):
            continue
        with open(packed, "rb") as f:
            blob = f.read()
        base = os.path.basename(prefix)
        d = out_dir if out_dir is not None else os.path.dirname(prefix) or "."
        os.makedirs(d, exist_ok=True)
        with open(idx_path) as f:
            header = f.readline().split()
            if not header or header[0] != "#sagephy-idx":
                raise ValueError(f"{idx_path}: not a sagephy index file")
            meta = dict(kv.split("=", 1) for kv in header[1:])
            first = first_index if first_index is not None else int(meta.get("first", 0))
            for i, line in enumerate(f):
                off, length, status = (int(x) for x in line.split("\t"))
                name = suffix
                if status != 0:
                    # a failed reference run leaves only its failure note: HostTreeGen.java:65 -> <prefix>.info, GuestTreeGen.java:50 -> <prefix>.pruned.info
                    if length == 0 or suffix != ".pruned.info":
                        continue
                    if meta.get("app") == "HostTreeGen":
                        name = ".info"
                with open(os.path.join(d, f"{base}_{first + i}{name}"), "wb") as out:
                    out.write(blob[off:off + length])
                written += 1
    return written


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("prefix")
    ap.add_argument("--first-index", type=int, default=None, help="default: the value recorded in the index")
    ap.add_argument("--out-dir", default=None)
    a = ap.parse_args()
    print(unpack(a.prefix, a.first_index, a.out_dir), "files written")
