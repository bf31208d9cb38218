"""Measure the FP64 roofline denominators on the GPU box (SURVEY.md section 6 asks for this).

Writes gpurun_out/fp64_peak.json: cuBLAS DGEMM (torch.matmul float64, burst best-of and sustained),
plus the DFMA / DMMA issue-rate micro-benchmark in tools/fp64_peak.cu.
Run on the box:  python tools/measure_fp64_peak.py
"""
import json
import os
import subprocess
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dgemm(n, reps):
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: back to back for ~3 s
    t0 = time.time()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    cnt = 0
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(4):
            torch.matmul(a, b)
        cnt += 4
        torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    sus = e0.elapsed_time(e1) / cnt
    fl = 2.0 * n ** 3
    return fl / best * 1e-9, fl / sus * 1e-9


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    for n in (4096, 8192):
        burst, sus = dgemm(n, 10)
        out[f"cublas_dgemm_{n}_tflops_burst"] = round(burst, 3)
        out[f"cublas_dgemm_{n}_tflops_sustained"] = round(sus, 3)
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    if os.path.exists(exe):
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        try:
            out["micro"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as ex:  # noqa: BLE001
            out["micro_error"] = f"{ex}: {r.stdout[-300:]} {r.stderr[-300:]}"
    q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                        "--format=csv,noheader"], capture_output=True, text=True)
    out["nvidia_smi_after"] = q.stdout.strip()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fp64_peak.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
