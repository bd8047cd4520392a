#!/usr/bin/env python
"""savsr_osa_prologue at the Vid4 shape (ci = 192, B = 17, 828 pooled partials) for ncu (bring-up aid):
    ncu --cache-control none --metrics gpu__time_duration.sum -k regex:osa_ -s 8 -c 4 python scripts/profile_osa.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402

for _ in range(3):
    G.check_osa_prologue(ci=192, B=17, npart=828, npix=25920)
