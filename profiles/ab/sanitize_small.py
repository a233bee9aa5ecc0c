"""Small end-to-end run for compute-sanitizer: 1c + 2c batched fits, AICc, resid stats, model cube on a 2 x 64 tile."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import _tiles as T
from mufasa_b200 import _abi, synth
from mufasa_b200.engine import get_engine
eng = get_engine(0)
for cfg in (2, 4):
    tile = T.sample_tile(cfg, nrows=2, xstep=4 if cfg == 2 else 8)
    spec, _ = T.pixel_major(tile, tile["guess1"])
    npix = spec.shape[0]
    d_spec = torch.from_numpy(spec).cuda()
    jobs = []
    for nc in (1, 2):
        lo, hi = T.limits_vectors(tile["fittype"], nc)
        g = synth.clamp_guesses(tile["guess%d" % nc], lo, hi)
        _, gg = T.pixel_major(tile, g)
        prob = _abi.make_problem(tile["fittype"], nc, tile["nchan"], npix, tile["v"][0], tile["dv"], lo, hi)
        jobs.append(dict(prob=prob, spectra=d_spec, guesses=torch.from_numpy(gg).cuda()))
    o1, o2 = eng.fit_batch(jobs)
    a = eng.aicc_global(jobs[0]["prob"], d_spec, o1["params"], o2["params"])
    r = eng.resid_stats(jobs[1]["prob"], d_spec, o2["params"])
    m = eng.model(jobs[1]["prob"], o2["params"])
    torch.cuda.synchronize()
    print("cfg", cfg, "npix", npix, "status hist", np.bincount(o2["status"].cpu().numpy().clip(0, 8), minlength=9).tolist(),
          "ncomp hist", np.bincount(a["ncomp"].cpu().numpy().astype(int), minlength=3).tolist())
