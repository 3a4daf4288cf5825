import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes
from splishsplash_b200.solver import build_b200_scene
from tests.parity import make_oracle, sync_state, scaled_err
sc = scenes.dam_break("small", dtype=np.float32)
ref, kind = make_oracle(sc, "f32"); dev = build_b200_scene(sc, "f32")
for s in range(8):
    sync_state(ref, dev); ref.step(1); dev.step(1)
    kr = ref.field_by_id("p / rho^2"); kd = dev.field("p / rho^2"); fr = ref.field_by_id("factor"); da = ref.field_by_id("advected density")
    i = np.argmax(np.abs(kr-kd))
    print(s, "max|k|", np.abs(kr).max(), "n(k>0)", (kr>0).sum(), "maxdiff", np.abs(kr-kd).max(), "at", i, "kr", kr[i], "kd", kd[i], "factor", fr[i], "dadv-1", da[i]-1, "dk/factor", np.abs(kr-kd)[i]/fr[i], "max dk/factor", (np.abs(kr-kd)/np.maximum(fr,1e-30)).max())
