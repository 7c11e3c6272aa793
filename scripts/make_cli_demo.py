"""Write a small noisy directory, a reference-format checkpoint (seeded default-init weights of the drop-in class) and the TOML
for the command-line demo (scripts/gpu_multi2.sh)."""
import os, sys
import torch
from scipy.io import wavfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fullsubnet-plus_b200")]
from fsnplus_b200.model import FullSubNet_Plus
from fsnplus_b200.synth import synth_clips
from fsnplus_b200.tools.inference import load_toml
out = sys.argv[1]
os.makedirs(os.path.join(out, "noisy"), exist_ok=True)
clips = synth_clips(24).numpy()
for i, c in enumerate(clips):
    wavfile.write(os.path.join(out, "noisy", f"clip{i:02d}.wav"), 16000, c)
for i in range(5):
    wavfile.write(os.path.join(out, "noisy", f"short{i}.wav"), 16000, clips[i][:32000])
toml_path = os.path.join(ROOT, "tests", "golden", "inference_reference.toml")
torch.manual_seed(0)
net = FullSubNet_Plus(**load_toml(toml_path)["model"]["args"])
torch.save({"model": net.state_dict(), "epoch": 58}, os.path.join(out, "ckpt.tar"))
open(os.path.join(out, "inference.toml"), "w").write(open(toml_path).read())
print("wrote", out)
