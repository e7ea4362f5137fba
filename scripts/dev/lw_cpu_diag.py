"""Development diagnostic (CPU only, run on the GPU box): where does the NumPy restatement leave the CPU emulation of
the kernels at 2152 x 2152?"""
import sys, os, ctypes as C, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
T = importlib.import_module("test_linear_winds_host")
from oracle import linear_winds_np as LW
so = "/tmp/liblw_kernels_host.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", os.path.join(ROOT, "tests/cpp/lw_kernels_host.cpp"), "-o", so])
emu = C.CDLL(so)
np.show_config() if "-v" in sys.argv else None
print("numpy", np.__version__, "cpu features:", [f for f in ("AVX2", "AVX512F", "AVX512_SKX", "AVX512_SPR") if np._core._multiarray_umath.__cpu_features__.get(f)])
n, waves, buffer, h0, U, V, Nsq = 2048, 64, 50, 100.0, 10.0, 0.0, 1e-4
x = np.arange(n, dtype=np.float64)[:, None]
terrain = (h0 * np.cos(2 * np.pi * waves * x / n) * np.ones((1, n))).astype(np.float32)
E = T.EmulatedLinearWinds(emu, np.ascontiguousarray(terrain.T), buffer, 2000.0)
tf, _ = LW.terrain_frequency(terrain, buffer)
print("tf diff", np.abs(E.tf.T - tf).max(), "max", np.abs(tf).max())
lt = LW.LinearTheoryData(E.fnx, E.fny, 2000.0)
print("k diff", np.abs(lt.k[:, 0] - E.kx).max(), "l diff", np.abs(lt.l[0] - E.ly).max())
f = np.float32
z = f(450.0)
uh_e, vh_e = E._spectrum(U, V, Nsq, z, f(100.0), 1)
# the restatement's uhat at one height, step by step
sig = (f(U) * lt.k + f(V) * lt.l).astype(f); sig[sig == 0] = LW.SMALL_VALUE
denom = (sig ** 2).astype(f).astype(np.complex128)
msq = (f(Nsq) / denom) * lt.kl
m = np.sqrt(msq); m = np.where(sig < 0, -m, m)
print("msq imag range", msq.imag.min(), msq.imag.max(), "m imag range", m.imag.min(), m.imag.max())
msq_r = (np.float64(f(Nsq)) / (sig.astype(np.float64) ** 2).astype(f).astype(np.float64)) * lt.kl.astype(np.float64)
print("msq real diff vs plain real arithmetic", np.abs(msq.real - msq_r).max() / np.abs(msq_r).max())
ph = m.real * np.float64(z)
e1 = np.exp(1j * m * np.float64(z))
e2 = np.cos(ph) + 1j * np.sin(ph)
print("exp(1j m z) vs cos + i sin:", np.abs(e1 - e2).max(), "max phase", np.abs(ph).max())
ineta = 1j * tf * e1
ineta = ineta / (lt.kl / ((0 - m) * sig))
uhat = LW.ifftshift2cc(lt.k * ineta)
print("uhat restatement vs emulation:", np.abs(uhat - uh_e.T).max(), "max", np.abs(uhat).max())
ineta2 = 1j * tf * e2 / (lt.kl / ((0 - m) * sig))
print("uhat with cos/sin instead of complex exp vs emulation:", np.abs(LW.ifftshift2cc(lt.k * ineta2) - uh_e.T).max())
