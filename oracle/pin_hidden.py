"""Pins oracle.decoder_forward(..., hidden=(h0, c0)) against the REAL reference LSTMDecoder.forward (src/student_model.py:205-256, the
`hidden` argument at :205 / :219-222 / :243) and writes tests/golden/hidden_init_case.pt = inputs + parameters + the reference's outputs
and gradients.  Runs only where /root/reference exists (the build container); the committed fixture is the pin on the GPU box.

    python oracle/pin_hidden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import kd_oracle as O  # noqa: E402
from oracle.pin_against_reference import load_reference, build_reference_student  # noqa: E402


def main():
    ref_sm, _ = load_reference()
    B, T, V, E, H, L, S = 5, 4, 60, 32, 64, 2, 49
    model = build_reference_student(ref_sm, V, E, H, L, seed=21, refinement=False)
    dec = model.decoder
    g = torch.Generator().manual_seed(22)
    feats = torch.randn(B, S, E, generator=g)
    cap = torch.randint(0, V, (T, B), generator=g)
    h0 = 0.5 * torch.randn(L, B, H, generator=g)
    c0 = 0.5 * torch.randn(L, B, H, generator=g)
    dout = torch.randn(T, B, V, generator=g)
    ok = True
    out = {}
    for dt, tol in ((torch.float64, 1e-12), (torch.float32, 2e-5)):
        d = dec.to(dt)
        for p_ in d.parameters():
            p_.grad = None
        f = feats.to(dt).clone().requires_grad_(True)
        outputs, hids, atts = d(f, cap, (h0.to(dt), c0.to(dt)))
        (outputs * dout.to(dt)).sum().backward()
        params = {"decoder." + k: v.detach().clone() for k, v in d.state_dict().items()}
        o2, h2, a2 = O.decoder_forward(params, feats.to(dt), cap, hidden=(h0.to(dt), c0.to(dt)))
        e = max(float((o2 - outputs).abs().max()), float((torch.stack(h2) - torch.stack(hids)).abs().max()),
                float((torch.stack(a2) - torch.stack(atts)).abs().max()))
        print(f"{str(dt)[6:]}: oracle vs reference with hidden=(h0, c0): max abs err {e:.2e}")
        ok &= e < tol
        if dt == torch.float32:
            out = {"meta": dict(B=B, T=T, V=V, E=E, H=H, L=L, S=S, torch=torch.__version__),
                   "params": {k: v.float() for k, v in params.items()}, "feats": feats, "captions": cap, "h0": h0, "c0": c0, "dout": dout,
                   "reference": {"outputs": outputs.detach(), "hidden_states": torch.stack(hids).detach(), "attention_weights": torch.stack(atts).detach(),
                                 "d_feats": f.grad.detach(), "grads": {"decoder." + k: v.grad.detach().clone() for k, v in d.named_parameters()}},
                   "generator": "oracle/pin_hidden.py (reference LSTMDecoder.forward with hidden)"}
    assert ok, "oracle does not match the reference"
    torch.save(out, os.path.join(ROOT, "tests", "golden", "hidden_init_case.pt"))
    print("ALL PINNED; wrote tests/golden/hidden_init_case.pt")


if __name__ == "__main__":
    main()
