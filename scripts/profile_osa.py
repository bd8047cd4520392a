import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import gpu_checks as G
for _ in range(3):
    G.check_osa_prologue(ci=192, B=17, npart=828, npix=25920)
