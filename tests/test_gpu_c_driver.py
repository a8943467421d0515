"""The C ABI used by a plain C program (examples/sed_infer.c: no Python, no tensor library) gives, bit for bit, the
results of the Python host engine on the same reference-layout parameters and waveforms."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from conftest import synthetic_sd
from sed_b200 import engine, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "examples", "sed_infer")
MT = "Cnn_9layers_Gru_FrameAtt"


def write_weights(path, sd):
    with open(path, "wb") as f:
        items = [(k, v) for k, v in sd.items() if v.dtype == torch.float32]
        f.write(struct.pack("<i", len(items)))
        for k, v in items:
            name = k.encode()
            f.write(struct.pack("<i", len(name)) + name)
            f.write(struct.pack("<i", v.dim()))
            f.write(struct.pack("<%dq" % v.dim(), *v.shape))
            f.write(v.contiguous().numpy().tobytes())


@pytest.mark.parametrize("sr,n_fft,hop,seconds,clips", [(16000, 512, 160, 10, 5), (32000, 1024, 320, 5, 3)])
def test_c_driver_equals_python_engine(tmp_path, sr, n_fft, hop, seconds, clips):
    assert os.path.isfile(DRIVER), "examples/sed_infer is not built: run __graft_entry__.build()"
    sd = synthetic_sd(MT, sr)
    wave = synth.synthetic_waveform(clips, sr * seconds, seed=77)
    wpath, xpath, opath = (str(tmp_path / n) for n in ("weights.bin", "wave.bin", "out.bin"))
    write_weights(wpath, sd)
    with open(xpath, "wb") as f:
        f.write(struct.pack("<ii", *wave.shape))
        f.write(wave.numpy().tobytes())
    r = subprocess.run([DRIVER, wpath, xpath, opath, str(n_fft), str(hop)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(opath, "rb").read()
    B, frames, classes = struct.unpack("<iii", raw[:12])
    clip = np.frombuffer(raw, np.float32, B * classes, 12).reshape(B, classes)
    frame = np.frombuffer(raw, np.float32, B * frames * classes, 12 + 4 * B * classes).reshape(B, frames, classes)
    pm = engine.PackedModel(sd, MT, n_fft, hop, torch.device("cuda:0"))
    out = pm.forward(wave.cuda())
    assert tuple(out["framewise_output"].shape) == (B, frames, classes)
    assert np.array_equal(out["clipwise_output"].cpu().numpy(), clip)
    assert np.array_equal(out["framewise_output"].cpu().numpy(), frame)
