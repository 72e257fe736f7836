import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import common, test_golden_gpu as tg
dev = torch.device("cuda:0")
g = np.load(os.path.join(ROOT, "tests/golden", sys.argv[1]))
a, dL = tg._args_from_golden(g, dev)
m = common.run_mine(a, dL)
P, W, H = int(g["P"]), int(g["W"]), int(g["H"])
ms = common.mine_sections(m, P, W, H)
pl = ms["point_list"].cpu().numpy(); ref = g["point_list"]
print("R", m["R"], int(g["R"]), "mismatch", (pl != ref).sum(), "first", np.argmax(pl != ref))
print("ranges mine", ms["ranges"].cpu().numpy()[:24].tolist())
print("ranges ref ", g["ranges"][:24].tolist())
