"""Dev tool: GPU timeline (torch.profiler/CUPTI) of the per-view work of the mapping step (C4):
prefilter -> decode -> rasterize -> loss -> backward -> bucket accumulate.  Busy vs span, top kernels, gaps."""
import os, sys, json, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from segs_slam_b200 import anchor_model, mapper
dev = torch.device("cuda:0")
NV = int(sys.argv[1]) if len(sys.argv) > 1 else 8
loss_kind = sys.argv[2] if len(sys.argv) > 2 else "default"
W, H, fx = 1200, 680, 600.0
tanx, tany = W / (2 * fx), H / (2 * fx)
model = anchor_model.synth_anchor_model(200_000, W, H, fx, fx, 1003, device=dev)
cams = anchor_model.circle_keyframes(64, 1.5, (0.0, 0.0, 3.25), tanx, tany, dev)[:NV]
g = torch.Generator(device="cpu").manual_seed(1)
target = (torch.rand(3, H, W, generator=g) * 0.5).to(dev)
if loss_kind == "fused":
    fm = mapper.FusedMapper(model, H, W, tanx, tany, torch.zeros(3, device=dev), lrs=1e-4)
    step = lambda: fm.step(cams, [target] * NV)
else:
    kw = {} if loss_kind == "default" else {"loss": loss_kind}
    render_loss = mapper.make_render_loss(model, cams, [target] * NV, H, W, tanx, tany, torch.zeros(3, device=dev), **kw)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    state = {"b": None}
    def step():
        loss, state["b"] = mapper.mapping_step(params, render_loss, NV, opt, state["b"])
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace_map.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
busy = sum(e["dur"] for e in ev)
print(f"views {NV} events {len(ev)} ({len(ev)/NV:.0f}/view) span {span/NV:.1f} us/view busy {busy/NV:.1f} us/view idle {(span-busy)/NV:.1f} us/view")
agg = {}
for e in ev:
    k = e["name"][:70]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += e["dur"]
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"  {d/NV:8.1f} us/view  n/view={n/NV:5.1f}  {k}")
gagg = {}
for x, y in zip(ev[:-1], ev[1:]):
    gp = y["ts"] - (x["ts"] + x["dur"])
    if gp > 0:
        gagg.setdefault((x["name"][:36], y["name"][:36]), [0, 0.0]); gagg[(x["name"][:36], y["name"][:36])][0] += 1; gagg[(x["name"][:36], y["name"][:36])][1] += gp
print("largest gap classes (us/view):")
for (x, y), (n, d) in sorted(gagg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {d/NV:7.1f}  n/view={n/NV:4.1f}  {x}  ->  {y}")
