import sys, os, math, torch
sys.path.insert(0, os.getcwd())
from sdvar_b200 import _cabi
from sdvar_b200.models.vqvae import _packed_w_up
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
B = 64
for (c, hw) in ((160, 128), (320, 64), (320, 32), (640, 16)):
    x = torch.randn(B, c, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(c, c, 3, padding=1).cuda()
    wpu = _packed_w_up(conv)
    wp = conv.weight.detach().permute(2, 3, 0, 1).reshape(9, c, c).bfloat16().contiguous()
    bias = conv.bias.detach().float().contiguous()
    y = torch.empty(B, c, 2 * hw, 2 * hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xu = torch.empty(B, c, 2 * hw, 2 * hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    t_f = timeit(lambda: _cabi.conv_up2x_nhwc(x, B, hw, hw, c, wpu, c, bias, y))
    def two():
        _cabi.upsample2x_nhwc(x, B, hw, hw, c, xu)
        _cabi.conv_nhwc(xu, B, 2 * hw, 2 * hw, c, wp, 9, c, bias, None, y=y)
    t_2 = timeit(two)
    print(f"up+conv {c}ch {hw}->{2*hw}: fused {t_f:.3f} ms   upsample+conv9 {t_2:.3f} ms")
