"""In-graph marginal cost of every step of the UNet evaluation: the step graph is re-captured WITHOUT one step at a time
(outputs are garbage, timing is not) and timed against the full graph on the same box.  Eager per-launch times
(bench.py --dump-layers) overstate small kernels: inside the graph, programmatic dependent launch hides launch latency and
prologues.  usage: ablate_steps.py [B H W] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import synthetic
from daclip_b200.unet import ConditionalUNet

B, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 256, 256)
REPS = int(sys.argv[4]) if len(sys.argv) >= 5 else 30
sd, kw = synthetic.unet_state_dict(0)
inp = {k: v.cuda() for k, v in synthetic.restoration_inputs(B, H, W, T=1, seed=3).items()}
net = ConditionalUNet(**kw)
net.load_state_dict(sd, strict=True)
net = net.cuda().eval()
eng = net.engine(B, H, W)
eng.set_inputs(inp["lq"], inp["lq"], inp["text_context"], inp["image_context"])
eng.set_time(37.0)
eng.replay()
torch.cuda.synchronize()
full_steps = list(eng.steps)


def capture(steps):
    eng.steps = steps
    eng.graph = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng._run_forked()
    return g


def timed(g, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


# every graph is captured first; the board then runs the full graph for a few seconds until its power / clock state has
# settled (the first second of a run is ~6 % faster than the steady state), and the ablated graphs are timed round-robin,
# each between two timings of the full graph
names = [n for n, _ in full_steps if n not in ("stem_input", "time_film")]
graphs = {}
for i, (name, _) in enumerate(full_steps):
    if name in names:
        graphs[name] = capture(full_steps[:i] + full_steps[i + 1:])
full = capture(full_steps)
for _ in range(600):
    full.replay()
torch.cuda.synchronize()
ROUNDS, R = 4, max(3, REPS // 6)
delta = {n: [] for n in names}
bases = []
for rnd in range(ROUNDS):
    b_prev = timed(full, R)
    for n in names:
        t = timed(graphs[n], R)
        b_next = timed(full, R)
        delta[n].append(0.5 * (b_prev + b_next) - t)
        bases.append(b_next)
        b_prev = b_next
bases.sort()
print(f"full graph: median {bases[len(bases) // 2]:.1f} us, min {bases[0]:.1f}, max {bases[-1]:.1f} ({len(bases)} timings of {R} replays)")
rows = []
for n in names:
    d = sorted(delta[n])
    med = 0.5 * (d[len(d) // 2 - 1] + d[len(d) // 2]) if len(d) % 2 == 0 else d[len(d) // 2]
    rows.append((med, n))
    print(f"{med:8.1f} us  (spread {d[0]:7.1f} .. {d[-1]:7.1f})  {n}", flush=True)
tot = sum(r[0] for r in rows)
print(f"sum of marginal costs: {tot:.1f} us")
cats = {}
for d, n in rows:
    key = ("attn1" if n.endswith("attn1") else "norm/aux" if any(k in n for k in ("norm", "fold", "prenorm")) else
           "linattn" if any(k in n for k in ("to_kv", "to_q", "to_out")) and "attn1" not in n else
           "xf_gemm" if any(k in n for k in ("proj_in", "proj_out", "attn1.", "ff.")) else "conv")
    cats[key] = cats.get(key, 0.0) + d
print({k: round(v, 1) for k, v in cats.items()})
