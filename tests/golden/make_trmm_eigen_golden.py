#!/usr/bin/env python3
"""Golden vectors for the TRMM post-processor (reference TRMM.cpp): the TRM and inverse_speed of the reference's
committed examples/infinite_GCR_TRMM/output.h5 and the eigen-pairs its TRMM.exe (Eigen::EigenSolver) wrote to
output_TRMM.h5 in the same directory.  Run in the build container (needs /root/reference); writes
tests/golden/trmm_eigen.npz."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import h5mini
D = "/root/reference/examples/infinite_GCR_TRMM/"
o, r = h5mini.File(D + "output.h5"), h5mini.File(D + "output_TRMM.h5")
np.savez(os.path.join(HERE, "trmm_eigen.npz"), TRM=o.root["TRM"].value, inverse_speed=o.root["inverse_speed"].value,
         **{k: r.root[k].value for k in ("alpha", "alpha_adj", "phi_mode", "phi_mode_adj")})
print({k: r.root[k].shape for k in ("alpha", "alpha_adj", "phi_mode", "phi_mode_adj")})
