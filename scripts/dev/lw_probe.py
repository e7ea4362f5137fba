"""Development probe (GPU box): where does the device's linear perturbation leave the restatement at large sizes?"""
import sys, os, ctypes as C, subprocess, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
import torch
import icar_b200
T = importlib.import_module("test_linear_winds_host")
so = "/tmp/liblw_kernels_host.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", os.path.join(ROOT, "tests/cpp/lw_kernels_host.cpp"), "-o", so])
emu = C.CDLL(so)
for n, waves in [(256, 8), (1000, 32), (1500, 48), (2048, 64)]:
    buffer, h0, U, V, Nsq = 50, 100.0, 10.0, 0.0, 1e-4
    x = np.arange(n, dtype=np.float64)[:, None]
    terrain = (h0 * np.cos(2 * np.pi * waves * x / n) * np.ones((1, n))).astype(np.float32)
    E = T.EmulatedLinearWinds(emu, np.ascontiguousarray(terrain.T), buffer, 2000.0)
    uhat, vhat = E._spectrum(U, V, Nsq, np.float32(450.0), np.float32(100.0), 4)
    E._inverse(uhat, vhat)
    up_cpu = E.up / 4
    # cuFFT (through torch) on the CPU-emulated spectrum
    tu = torch.fft.ifft2(torch.from_numpy(uhat).cuda()) * uhat.size
    d_fft = float(np.abs(tu.cpu().numpy() / 4 - up_cpu).max())
    dom = icar_b200.Domain(icar_b200.Grid.single_image(n, 2, n), icar_b200.Options(dx=2000.0))
    dom.linear_winds_init(terrain.T, buffer=buffer)
    u, v = dom.linear_perturbation(U, V, Nsq, 400.0, 800.0, minimum_step=100.0)
    d_dev = float(np.abs(u - up_cpu).max())
    tfd = float(np.abs(dom.terrain_frequency() - E.tf).max())
    print(n, E.fnx, "max|u|", float(np.abs(up_cpu).max()), "cuFFT(torch) on emulated spectrum:", d_fft, "device path:", d_dev, "tf:", tfd, flush=True)
    dom.close()
