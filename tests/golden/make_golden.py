"""Copies the reference's own golden vectors for the harmonic test into tests/golden/
(they cannot be read from /root/reference on the GPU box). Run in the build container:
    python tests/golden/make_golden.py
Sources: /root/reference/test/harmonic/harm.r0x, harm.E0x (x and E_x of particle 0 for the
1200 iterations of conf/harmonic.conf, printed by test/harmonic.c:112-120) and E.csv (a
ParaView line probe at y=4 of the 64x64 variant: E_X, E_Y, phi, rho)."""
import os
import shutil

REF = "/root/reference/test/harmonic"
HERE = os.path.dirname(os.path.abspath(__file__))
for f in ("harm.r0x", "harm.E0x", "E.csv"):
    shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, f))
    print("copied", f)
