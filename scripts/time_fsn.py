import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fullsubnet-plus_b200")]
from fsnplus_b200.model import Model, FullSubNet_Plus
import bench
torch.manual_seed(0)
dev = "cuda:0"
cfg = dict(sb_num_neighbors=15, fb_num_neighbors=0, num_freqs=257, look_ahead=2, sequence_model="LSTM", fb_output_activate_function="ReLU",
           sb_output_activate_function=False, fb_model_hidden_size=512, sb_model_hidden_size=384, weight_init=False)
def timeit(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
with torch.no_grad():
    m = Model(**cfg).to(dev).eval()
    for B in (1, 8, 64):
        x = torch.rand(B, 1, 257, 188, device=dev)
        print(f"fullsubnet.Model B={B}: {timeit(lambda: m(x)):.2f} ms  (sub-band LSTM {m.last_lstm_ms():.2f} ms)")
    p = FullSubNet_Plus(**bench.default_cfg()).to(dev).eval()
    for B in (1, 8):
        x = torch.rand(B, 1, 257, 188, device=dev)
        print(f"FullSubNet_Plus B={B}: {timeit(lambda: p(x, x - 0.5, x - 0.3)):.2f} ms  (sub-band LSTM {p.last_lstm_ms():.2f} ms)")
