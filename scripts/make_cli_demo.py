"""Write a small noisy directory, a reference-format checkpoint (seeded default-config weights) and the TOML for the CLI demo."""
import os, sys
import numpy as np, torch
from scipy.io import wavfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fullsubnet-plus_b200")]
from oracle import fsn_oracle as O
out = sys.argv[1]
os.makedirs(os.path.join(out, "noisy"), exist_ok=True)
clips = O.synth_clips(24).astype(np.float32)
for i, c in enumerate(clips):
    wavfile.write(os.path.join(out, "noisy", f"clip{i:02d}.wav"), 16000, c)
for i in range(5):
    wavfile.write(os.path.join(out, "noisy", f"short{i}.wav"), 16000, clips[i][:32000])
cfg = O.default_plus_config()
torch.save({"model": {k: torch.from_numpy(v) for k, v in O.make_params_plus(cfg, seed=0).items()}, "epoch": 58}, os.path.join(out, "ckpt.tar"))
open(os.path.join(out, "inference.toml"), "w").write(open(os.path.join(ROOT, "tests", "golden", "inference_reference.toml")).read())
print("wrote", out)
