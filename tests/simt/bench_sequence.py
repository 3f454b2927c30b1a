"""simt-check: the sequence of C-ABI calls bench.py makes (device initialiser, run, run_timed, stage
timers, the staged stages, image upload / step / download with the four grids, the drop-in's two
coherence modes), at a small size against the interpreted kernels. Run by tests/test_simt_check.py;
it guards the bench's use of the API where no GPU exists. Test infrastructure only."""
import os
import sys
import ctypes as C

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import numpy as np
from cpic_b200 import Sim, load_conf

conf = os.path.join(ROOT, "conf", "2d-2species-small.conf")
params, run = load_conf(conf)
sim = Sim(params)
nps=10000
for i in range(2):
    sim.init_uniform(i, nps, id0=0, vx=[5.0,3.0][i], vy=0.0, seed=138+i)
sim.pre_step(); sim.sync()
sim.run(3)
sim.timing(False)
ms = sim.run_timed(5); print("run_timed", ms)
_, launches0 = sim.get_timing(); print("launches", launches0)
sim.timing(True); sim.run(5); stage_ms, launches = sim.get_timing(); print(stage_ms, launches)
sim.timing(False)
sim.step_staged(); sim.sync(); sim.timing(True)
for _ in range(3): sim.step_staged()
sim.sync(); st,_=sim.get_timing(); print(st); sim.timing(False)
L=sim.L
nbytes=L.cpic_b200_image_bytes(sim.h); host=L.cpic_b200_host_alloc(nbytes)
fields={k: np.empty(sim.field_shape(k)) for k in ("rho","phi","Ex","Ey")}
assert L.cpic_b200_image_download(sim.h, host, nbytes)==0
for _ in range(3):
    assert L.cpic_b200_image_upload(sim.h, host, nbytes)==0
    sim.step()
    assert L.cpic_b200_image_download(sim.h, host, nbytes)==0
    for k,a in fields.items():
        assert L.cpic_b200_get_field(sim.h, {"rho":0,"phi":1,"Ex":2,"Ey":3}[k], a.ctypes.data_as(C.c_void_p))==0
sim.sync()
for wp in (True, False):
    for _ in range(3):
        sim.step()
        if wp: assert L.cpic_b200_image_download(sim.h, host, nbytes)==0
        for k,a in fields.items():
            assert L.cpic_b200_get_field(sim.h, {"rho":0,"phi":1,"Ex":2,"Ey":3}[k], a.ctypes.data_as(C.c_void_p))==0
    sim.sync()
L.cpic_b200_host_free(host)
print("n", sim.num_particles(0), sim.num_particles(1), "E", sim.energy() if hasattr(sim,'energy') else None)
sim.close(); print("DRY RUN OK")
