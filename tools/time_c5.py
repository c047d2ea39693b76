"""Config-5 block assembly timing (one line); used with MADB_LIB variants."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import mfem_ad_b200 as M
import bench_configs as BC
ctx = M.Context(0)
BC.config5(ctx, int(sys.argv[1]) if len(sys.argv) > 1 else 1)
