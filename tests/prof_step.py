"""Timing of ModelEnv.step at MBPO scale (config 4): device-resident tensors and numpy in/out (reference semantics)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from mbrl_lib_b200 import synthetic as syn
from test_gpu_parity import make_env, DEV

for precision in ("bf16_tc", "f32"):
    spec, arrays, env = make_env("mbpo_halfcheetah", precision, ts1="tile_shuffle")
    B = 100000
    inp = syn.make_step_inputs(spec, B)
    act = torch.from_numpy(inp["act"]).to(DEV)
    state = env.reset(inp["obs"], return_as_np=False)
    for _ in range(3):
        env.step(act, state, sample=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        env.step(act, state, sample=True)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    st_np = env.reset(inp["obs"], return_as_np=True)
    t0 = time.perf_counter()
    for _ in range(5):
        st_np = env.reset(inp["obs"], return_as_np=True)
        env.step(inp["act"], st_np, sample=True)
    dt = (time.perf_counter() - t0) / 5
    print(f"{precision}: device-resident step {ms:.3f} ms -> {B/ms*1e3:.3e} transitions/s ; reset+step numpy in/out {dt*1e3:.2f} ms -> {B/dt:.3e} transitions/s")
