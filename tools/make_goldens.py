#!/usr/bin/env python3
"""Regenerates tests/golden/species100.pruned.tree with the CPU oracle (the Java reference cannot run in this image)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_util import oracle_run
f = oracle_run("HostTreeGen", ["-s", "7", "-min", "100", "-max", "100", "1.0", "5.0", "0.05", "species100"])["files"]
open(os.path.join(ROOT, "tests", "golden", "species100.pruned.tree"), "wb").write(f["species100.pruned.tree"])
