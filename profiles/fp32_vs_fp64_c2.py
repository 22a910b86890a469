import sys, torch, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import fno_oracle as O
torch.set_num_threads(16)
s = (20, 256, 512, 3)
for gain in (1.0, 300.0):
    torch.manual_seed(41)
    sd = O.init_state(2, (12, 16), 4, 64, s, s)
    O.randomize_bn(sd, 42)
    sd = {k: (v * gain if k.startswith("spectral_convs.") else v.clone()) for k, v in sd.items()}
    torch.manual_seed(9)
    x = torch.randn(1, *s)
    sd64 = {k: (v.double() if v.is_floating_point() else (v.to(torch.cdouble) if v.is_complex() else v)) for k, v in sd.items()}
    with torch.no_grad():
        t0 = time.time()
        y32 = O.fno2d_forward(sd, x, s)
        y64 = O.fno2d_forward(sd64, x.double(), s)
    print("gain", gain, "fp32 oracle vs fp64 oracle rel_l2 =", O.rel_l2(y32, y64), f"({time.time()-t0:.1f}s)")
