"""Golden vectors of the TRAINING-side adapter geometry produced by the reference's own
class: imports /root/reference/MLGWSC-1/train.py (third-party modules stubbed, oracle QScan injected for ml4gw)
and runs its `QTransformAdapter` (128x128 Q-spectrogram, 32/64/128 CNN; train.py:78-160) on seeded strain.
    python tests/golden/make_train_adapter_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_reference_train():
    from oracle import qscan as OQ
    stubs = {}
    for name in ("h5py", "peft", "ml4gw", "ml4gw.transforms"):
        if name not in sys.modules:
            stubs[name] = types.ModuleType(name)
    if "h5py" in stubs:
        stubs["h5py"].File = object
    if "peft" in stubs:
        stubs["peft"].LoraConfig = object
        stubs["peft"].get_peft_model = lambda *a, **k: None
    if "ml4gw.transforms" in stubs:
        stubs["ml4gw.transforms"].QScan = OQ.QScan
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("ref_train_golden", os.path.join(REF, "MLGWSC-1/train.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_train_golden"] = mod          # @dataclass needs the module registered while it executes
        spec.loader.exec_module(mod)
    finally:
        for k in stubs:
            sys.modules.pop(k, None)
    return mod


def main():
    ref = load_reference_train()
    torch.manual_seed(23)
    adapter = ref.QTransformAdapter(n_detectors=2).eval()
    with torch.no_grad():
        adapter.scale.fill_(0.8)
        adapter.bias.fill_(-0.2)
        adapter.film_gamma.copy_(torch.tensor([1.1, 0.7]))
        adapter.film_beta.copy_(torch.tensor([0.0, 0.3]))
    g = torch.Generator().manual_seed(19)
    x = torch.randn(3, 2, 2048, generator=g)
    t = torch.arange(2048) / 2048.0
    x[1, 0] += 6.0 * torch.exp(-(t - 0.6) ** 2 / (2 * 0.02 ** 2)) * torch.sin(2 * np.pi * 250.0 * t)
    with torch.no_grad():
        feats = adapter(x)
    out = {"strain": x.numpy(), "feats_sub": feats[..., ::25].numpy(), "feats_rows": np.array([0, 41, 79]),
           "feats_fullrows": feats[:, :, [0, 41, 79], :].numpy()}
    for k, v in adapter.state_dict().items():
        if not k.startswith("q_transform."):
            out["adapter." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "train_adapter_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
