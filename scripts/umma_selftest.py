import ctypes, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distegnn_b200 import _lib
lib = _lib.load()
fn = lib.distegnn_selftest_umma
fn.argtypes = [ctypes.c_void_p]*3 + [ctypes.c_int, ctypes.c_void_p]
fn.restype = ctypes.c_int
torch.manual_seed(0)
dev = torch.device('cuda:0')
A = torch.randn(128, 64, device=dev)
W = torch.randn(64, 64, device=dev) / 8
ref = (A.double() @ W.double().t())
ref32 = A @ W.t()
print("fp32 torch vs fp64:", float((ref32.double()-ref).abs().max()))
variants = [int(v) for v in sys.argv[1:]] or [0]
for variant in variants:
    D = torch.zeros(128, 64, device=dev)
    rc = fn(A.data_ptr(), W.data_ptr(), D.data_ptr(), variant, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = float((D.double()-ref).abs().max())
    print(f"variant {variant}: rc={rc} max err {err:.3e}  (ref scale {float(ref.abs().max()):.2f})", flush=True)
