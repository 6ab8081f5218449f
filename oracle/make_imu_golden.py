"""Pin oracle/imu_encoder.py on the reference's own InertialEncoder and freeze tests/golden/imu_encoder.pt.

Run HERE (authoring container, /root/reference mounted):   python -m oracle.make_imu_golden
The fixture stores the seed, the input and the REFERENCE class's output (the weights are regenerated from the seed:
same construction order -> same RNG draws; 2.9 MB of projection weights stay out of the repo)."""

import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def main():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    from src.models.Encoder import InertialEncoder as RefEncoder          # the reference's own class
    from oracle.imu_encoder import OracleInertialEncoder, imu_like, randomize_batchnorm
    opt = SimpleNamespace(seq_len=11, imu_dropout=0.0, i_f_len=256)
    torch.manual_seed(0)
    ref = RefEncoder(opt)
    randomize_batchnorm(ref, seed=0)
    ref.eval()
    torch.manual_seed(0)
    ora = OracleInertialEncoder(opt)
    randomize_batchnorm(ora, seed=0)
    ora.eval()
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), ora.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), f"state_dict differs at {ka}"
    x = imu_like(3, 10, seed=1)
    with torch.no_grad():
        y_ref = ref(x)
        y_ora = ora(x)
    assert y_ref.shape == (3, 10, 256)
    assert torch.equal(y_ref, y_ora), f"restatement differs from the reference: {(y_ref - y_ora).abs().max().item():.3e}"
    out = os.path.join(ROOT, "tests", "golden", "imu_encoder.pt")
    torch.save(dict(seed=0, bn_seed=0, x=x, y=y_ref, i_f_len=256, seq_len=11,
                    proj_w_checksum=ref.proj.weight.double().sum().item(),
                    conv3_w_checksum=ref.encoder_conv[8].weight.double().sum().item()), out)
    print("reference InertialEncoder == oracle restatement (bit-identical on CPU); wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
