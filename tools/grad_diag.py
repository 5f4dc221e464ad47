import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import spe_oracle as O
from spe_b200 import factory
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_xxs24_224"
gold = torch.load(f"tests/golden/{name}.pt", weights_only=False)
m = gold["meta"]; cfg = O.SPEConfig(**m["cfg"])
params = O.make_params(cfg, m["seed"])
images, targets = O.make_inputs(cfg, m["batch"], m["height"], m["width"], seed=m["seed"], max_gt=m["max_gt"], repeat=m["repeat"])
dev = torch.device("cuda")
model = factory.build_detector(cfg, dev); model.load_state_dict(params); model.train()
crit = factory.build_criterion(cfg, m["losses"], gamma=m["gamma"], device=dev); crit.eval()
out = model(images.to(dev))
ld = crit(out[0], [{k: v.to(dev) for k, v in t.items()} for t in targets])
loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict); loss.backward()
_, _, _, og, _ = O.train_step(params, cfg, images, targets, m["losses"], gamma=m["gamma"])
rows = []
for k, p in model.named_parameters():
    g = (p.grad if p.grad is not None else torch.zeros_like(p)).float().cpu()
    rows.append(((g - og[k]).norm().item() ** 2, og[k].norm().item(), k))
tot = sum(r[0] for r in rows); den = sum(r[1] ** 2 for r in rows)
print("aggregate", (tot / den) ** 0.5)
for e, n, k in sorted(rows, reverse=True)[:14]:
    print("%-60s err2 share %.3f  rel %.4f  |g| %.4f" % (k, e / tot, e ** 0.5 / (n + 1e-12), n))
