"""One full enhance() call with N=1 PC step (2 network evaluations + sampler/STFT kernels) for B=8 inside a
cudaProfilerStart/Stop window (ncu --profile-from-start off ...)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from storm_b200 import ScoreModel, ncsnpp_spec  # noqa: E402
from oracle.storm_oracle import synth_utterance  # noqa: E402

B = 8
model = ScoreModel(backbone="ncsnpp", sde="ouve", init_scale=1.0)
model.dnn.load_state_dict(ncsnpp_spec.random_state_dict(model.dnn.cfg, seed=0, init_scale=1.0))
model.cuda()
y = torch.stack([synth_utterance(b, 32640) for b in range(B)]).cuda()
sc = model._sampler_cfg("reverse_diffusion", "ald", 1, 1, 0.5)
sp = model.data_module.spec_config()
eng = model.dnn.engine()
for _ in range(2):
    eng.enhance(y, sc, sp, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.enhance(y, sc, sp, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
