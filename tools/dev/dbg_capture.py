import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import spe_oracle as O
from spe_b200 import factory
from spe_b200.engine import TrainStep
dev = torch.device("cuda")
cfg = O.tiny_config()
model = factory.build_detector(cfg, dev).train()
model.load_state_dict(O.make_params(cfg, 5))
mode = sys.argv[1]
train = mode.endswith("train")
crit = factory.build_criterion(cfg, device=dev, match_ratio=5)
crit_r = factory.build_criterion(cfg, refine=True, device=dev, match_ratio=5)
(crit.train(), crit_r.train()) if train else (crit.eval(), crit_r.eval())
images, targets = O.make_inputs(cfg, 2, 48, 64, seed=5, max_gt=3)
tgd = [{k: v.to(dev) for k, v in t.items()} for t in targets]
trd = [dict(t, scores=torch.full((len(t["labels"]),), 0.6, device=dev)) for t in tgd]
order = (False, True) if mode.startswith("both") else (True,)
for graph in order:
    step = TrainStep(model, crit, crit_r, graph=graph, max_gt=20)
    for _ in range(2):
        loss = step(images.to(dev), tgd, trd)[0]
    print(mode, "graph", graph, float(loss), flush=True)
