import ctypes, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.twin_backend import load_testing
lib = load_testing()
fn = lib.distegnn_selftest_umma
torch.manual_seed(0)
dev = torch.device('cuda:0')
scale = float(os.environ.get("ASCALE", "1"))
A = torch.randn(128, 64, device=dev) * scale
W = torch.randn(64, 64, device=dev) / 8
ref = (A.double() @ W.double().t())
print("fp32 torch vs fp64:", float(((A @ W.t()).double()-ref).abs().max()), "ref scale", float(ref.abs().max()))
for variant in [int(v) for v in sys.argv[1:]] or [0]:
    D = torch.zeros(128, 64, device=dev)
    rc = fn(A.data_ptr(), W.data_ptr(), D.data_ptr(), variant, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    e = (D.double()-ref).abs()
    print(f"variant {variant}: rc={rc} max err {float(e.max()):.3e} rms {float(e.pow(2).mean().sqrt()):.3e}", flush=True)
