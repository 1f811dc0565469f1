import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "cater_b2.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_weights(golden):
    from textocvp_b200 import weights
    m = golden["meta"]
    savi_sd = weights.savi_state_dict(m["savi_seed"], bias_scale=m["bias_scale"], ln_jitter=m["ln_jitter"])
    pred_sd = weights.predictor_state_dict(m["pred_seed"], mlp_out_scale=m["mlp_out_scale"],
                                           ln_jitter=m["ln_jitter"])
    videos, text, noise = weights.synthetic_inputs(m["B"], m["T"], m["L"], seed=m["input_seed"])
    init = savi_sd["initializer.slots_mu"] + savi_sd["initializer.slots_sigma"] * noise
    return dict(savi_sd=savi_sd, pred_sd=pred_sd, videos=videos, text=text, init=init)


@pytest.fixture(scope="session")
def golden_dino():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "cliport_b2.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_dino_weights(golden_dino):
    from textocvp_b200 import weights
    m = golden_dino["meta"]
    dsd = weights.dino_state_dict(m["dino_seed"], img_size=m["img_size"], num_patches=m["N"],
                                  bias_scale=m["bias_scale"], ln_jitter=m["ln_jitter"], bn_jitter=m["bn_jitter"])
    psd = weights.predictor_state_dict(m["pred_seed"], mlp_out_scale=m["mlp_out_scale"], ln_jitter=m["ln_jitter"])
    feats, text, noise = weights.synthetic_dino_inputs(m["B"], m["T"], m["N"], L=m["L"], seed=m["input_seed"])
    init = dsd["initializer.slots_mu"] + dsd["initializer.slots_sigma"] * noise
    return dict(dino_sd=dsd, pred_sd=psd, feats=feats, text=text, init=init)


def pytest_sessionfinish(session, exitstatus):
    """Measured errors of the parity tests next to their gates (oracle/parity_log.py) -> gpurun_out/parity_r2.md."""
    try:
        from oracle import parity_log
        if parity_log.RECORDS:
            parity_log.dump(os.path.join(ROOT, "gpurun_out", "parity_r2.md"))
    except Exception as e:                       # the log must never turn a green run red
        print(f"parity log not written: {e}")
