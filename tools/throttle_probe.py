"""How does the step time evolve under sustained load?  Times consecutive chunks of forwards with CUDA events and samples
nvidia-smi clocks/power next to them (decides how many steps a bench needs before it reports a sustained number)."""
import argparse, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=30)
ap.add_argument("--chunk", type=int, default=10)
ap.add_argument("--head", default="sls")
ap.add_argument("--mode", default="device", choices=["device", "submit", "host"])
a = ap.parse_args()
m = (sls_b200.ModelSLS if a.head == "sls" else sls_b200.Model)(None, "cuda", cp_path=None).to("cuda").eval()
eng = m.engine()
head = sls_b200.HEAD_SLS if a.head == "sls" else sls_b200.HEAD_SAE
wav = [eng.synth_clips(i * 64, 64) for i in range(4)]
host = [w.cpu().pin_memory() for w in wav]
outs = [torch.empty(64, dtype=torch.float32, pin_memory=True) for _ in range(4)]
for i in range(3):
    eng.forward(wav[i % 4], head, sls_b200.PREC_BF16)
torch.cuda.synchronize()
q = "clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"
for c in range(a.chunks):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(a.chunk):
        if a.mode == "device":
            eng.forward(wav[i % 4], head, sls_b200.PREC_BF16)
        elif a.mode == "submit":
            eng.score_submit(host[i % 4], head, sls_b200.PREC_BF16, out=outs[i % 4])
        else:
            eng.score_host(host[i % 4], head, sls_b200.PREC_BF16)
    t_enq = time.perf_counter() - t0
    if a.mode == "submit":
        eng.score_wait()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    smi = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
    print(f"chunk {c:3d} [{a.mode}]: {e0.elapsed_time(e1) / a.chunk:7.3f} ms/step (events)  wall {wall / a.chunk * 1e3:7.3f}  cpu-enqueue {t_enq / a.chunk * 1e3:6.3f} ms/step | smi(after) {smi}", flush=True)
