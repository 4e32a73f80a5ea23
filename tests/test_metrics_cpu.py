"""utils/calculate_metrics.py (reference utils/calculate_metrics.py:78-224): known-answer checks of the
metrics that need no pretrained network, and the CLI flow on a pair of small PNG directories."""
import math

import numpy as np
import pytest
import torch


def test_psnr_mse_known_answers():
    from cfgpp_b200.utils import calculate_metrics as M
    a = np.zeros((1, 3, 8, 8), np.float32)
    b = np.full((1, 3, 8, 8), 16.0, np.float32)
    assert M.mean_squared_error(a, b) == 256.0
    assert M.peak_signal_noise_ratio(a, b, 255.0) == pytest.approx(10 * math.log10(255.0 ** 2 / 256.0))
    assert M.peak_signal_noise_ratio(a, a) == float("inf")


def test_frechet_distance_known_answers():
    from cfgpp_b200.utils import calculate_metrics as M
    rng = np.random.default_rng(0)
    f = rng.normal(size=(500, 6))
    assert abs(M.FID.from_features(f, f)) < 1e-8                       # identical sets
    mu, s = np.zeros(3), np.eye(3)
    assert M.frechet_distance(mu, s, mu + 2.0, s) == pytest.approx(12.0)          # ||dmu||^2 = 3 * 4
    assert M.frechet_distance(mu, s, mu, 4 * s) == pytest.approx(3 * (1 + 4 - 2 * 2))  # Tr(S1 + S2 - 2 sqrt(S1 S2))


def test_mnc_of_a_kernel_with_itself_is_one():
    from cfgpp_b200.utils import calculate_metrics as M
    k = torch.zeros(1, 1, 9, 9)
    k[0, 0, 4, 2:7] = 0.2
    assert M.MNC.calculate_mnc(k, k).item() == pytest.approx(1.0, abs=1e-5)


def test_cli_flow_on_png_directories(tmp_path):
    from PIL import Image
    from cfgpp_b200.utils import calculate_metrics as M
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    rng = np.random.default_rng(1)
    for i in range(3):
        img = rng.integers(0, 256, size=(16, 16, 3), dtype=np.uint8)
        Image.fromarray(img).save(tmp_path / "a" / f"{i}.png")
        Image.fromarray(np.clip(img.astype(int) + 4, 0, 255).astype(np.uint8)).save(tmp_path / "b" / f"{i}.png")
    out = M.run(tmp_path / "a", tmp_path / "b", "unit", log_path=str(tmp_path / "r.log"), metrics=("PSNR", "MSE", "LPIPS", "FID"))
    mean, std = out["PSNR"]
    assert 35.0 < mean < 37.0 and std >= 0           # +4 grey levels (a little clipping) -> ~36 dB
    assert out["MSE"][0] == pytest.approx((4 / 255) ** 2, rel=0.1)
    assert out["LPIPS"] is None and out["FID"] is None  # optional packages / weights are not installed offline
    assert "Metric Calculation for unit" in (tmp_path / "r.log").read_text()
