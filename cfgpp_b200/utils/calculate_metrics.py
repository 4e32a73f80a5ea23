"""Image-quality metrics of the evaluation flow (reference `utils/calculate_metrics.py:78-224`, CLI :206-224):

    python -m cfgpp_b200.utils.calculate_metrics --input_dir gen/ --label_dir ref/ --exp_name run1 [--log_path result.log] [--gpu 0]

Pairs the sorted `*.png` of the two directories and reports mean / std of each paired metric, then FID over the two
sets — same flags, same logger output shape, same metric classes (`Metric` base, `MSE`, `PSNR`, `LPIPS`, `FID`, `MNC`).
The reference gets PSNR / MSE from scikit-image, LPIPS from the `lpips` package and FID from `pytorch_fid`; none of them
(nor their pretrained VGG / Inception weights) can be installed offline, so PSNR / MSE / MNC and the Frechet distance
itself are implemented here directly, and LPIPS / FID use the optional packages when they are importable and otherwise
report `unavailable` instead of a number (they never invent one). `FID.from_features` computes the distance from any
caller-provided feature matrices. Outside the hot path: host-side numpy / torch, no CUDA kernels."""
from __future__ import annotations

import argparse
import logging
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def prepare_logger(log_path: str):
    logger = logging.getLogger("Metric")
    logger.setLevel(logging.INFO)
    if not logger.handlers:
        stream = logging.StreamHandler()
        stream.setFormatter(logging.Formatter("%(asctime)s-%(levelname)s >> %(message)s"))
        logger.addHandler(stream)
        logger.addHandler(logging.FileHandler(log_path))
    return logger


def load_rgb(path: Path) -> torch.Tensor:
    """PNG -> float tensor (1,3,H,W) in [0,1] (what torchvision's ToTensor yields in the reference's dataset)."""
    from PIL import Image
    arr = np.asarray(Image.open(path).convert("RGB"), dtype=np.float32) / 255.0
    return torch.from_numpy(arr).permute(2, 0, 1)[None]


def mean_squared_error(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))


def peak_signal_noise_ratio(label: np.ndarray, test: np.ndarray, data_range: float = 255.0) -> float:
    """10 log10(data_range^2 / MSE) — skimage.metrics.peak_signal_noise_ratio semantics (inf for identical images)."""
    err = mean_squared_error(label, test)
    return float("inf") if err == 0 else float(10.0 * np.log10(data_range ** 2 / err))


def frechet_distance(mu1: np.ndarray, sigma1: np.ndarray, mu2: np.ndarray, sigma2: np.ndarray, eps: float = 1e-6) -> float:
    """||mu1 - mu2||^2 + Tr(S1 + S2 - 2 (S1 S2)^(1/2)) — the FID formula (Heusel et al. 2017)."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1).astype(np.float64), np.atleast_1d(mu2).astype(np.float64)
    sigma1, sigma2 = np.atleast_2d(sigma1).astype(np.float64), np.atleast_2d(sigma2).astype(np.float64)
    diff = mu1 - mu2
    covmean = np.asarray(linalg.sqrtm(sigma1.dot(sigma2)))
    if not np.isfinite(covmean).all():  # nearly singular product: regularise, as pytorch_fid does
        off = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + off).dot(sigma2 + off))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2.0 * np.trace(covmean))


class Metric:
    paired = True

    def __init__(self, input_dir, label_dir, logger, device):
        self.input_dir, self.label_dir = Path(input_dir), Path(label_dir)
        self.logger, self.device = logger, device

    def retrieve_img_paths(self, directory: Path):
        return sorted(directory.glob("*.png"))

    def preprocessing(self, img: torch.Tensor):
        return img

    def metric_fn(self, label, img):
        raise NotImplementedError

    def compute(self) -> Optional[Tuple[float, float]]:
        self.logger.info(f"Start to calculate metric {self}.")
        ins, labels = self.retrieve_img_paths(self.input_dir), self.retrieve_img_paths(self.label_dir)
        assert len(ins) == len(labels), (f"Two file lists should have the same number of files. "
                                          f"Got {len(ins)} and {len(labels)}")
        vals = []
        with torch.no_grad():
            for pi, pl in zip(ins, labels):
                a, b = self.preprocessing(load_rgb(pi)), self.preprocessing(load_rgb(pl))
                if isinstance(a, torch.Tensor):
                    a, b = a.to(self.device), b.to(self.device)
                vals.append(float(self.metric_fn(b, a)))
        t = torch.tensor(vals, dtype=torch.float64)
        mean, std = t.mean().item(), (t.std().item() if len(vals) > 1 else float("nan"))
        self.logger.info(f"Result: mean={mean}  std={std}")
        return mean, std


class MSE(Metric):
    def __str__(self):
        return "MSE"

    def preprocessing(self, img):
        return img.numpy()

    def metric_fn(self, label, img):
        return mean_squared_error(label, img)


class PSNR(Metric):
    def __str__(self):
        return "PSNR"

    def preprocessing(self, img):
        return img.numpy() * 255

    def metric_fn(self, label, img):
        return peak_signal_noise_ratio(label, img, data_range=255.0)


class LPIPS(Metric):
    """Needs the `lpips` package and its pretrained VGG / AlexNet weights (reference :133-141)."""

    def __init__(self, input_dir, label_dir, logger, device, net: Optional[str] = "vgg"):
        super().__init__(input_dir, label_dir, logger, device)
        try:
            import lpips  # noqa: F401
            self.model = lpips.LPIPS(net=net).to(device)
        except Exception as e:  # noqa: BLE001 — not installable offline
            self.model, self.why = None, repr(e)

    def __str__(self):
        return "LPIPS"

    def metric_fn(self, label, img):
        return self.model(label, img).item()

    def compute(self):
        if self.model is None:
            self.logger.info(f"Result: LPIPS unavailable ({self.why})")
            return None
        return super().compute()


class FID(Metric):
    """Frechet Inception distance between the two directories (reference :167-183, dims=2048, batch_size=1)."""
    paired = False

    def __str__(self):
        return "FID"

    @staticmethod
    def from_features(f1: np.ndarray, f2: np.ndarray) -> float:
        """FID from two (n, d) feature matrices (Inception pool3 activations in the standard protocol)."""
        return frechet_distance(f1.mean(0), np.cov(f1, rowvar=False), f2.mean(0), np.cov(f2, rowvar=False))

    def compute(self):
        self.logger.info(f"Start to calculate metric {self}.")
        try:
            from pytorch_fid.fid_score import calculate_fid_given_paths
        except Exception as e:  # noqa: BLE001 — pytorch_fid and its Inception weights are not available offline
            self.logger.info(f"Result: FID unavailable ({e!r}); FID.from_features() accepts precomputed features")
            return None
        value = calculate_fid_given_paths([str(self.input_dir), str(self.label_dir)], batch_size=1, device=self.device,
                                          dims=2048)
        self.logger.info(f"Result: {value}")
        return value


class MNC(Metric):
    """Maximum of normalized convolution (Hu & Yang, ECCV 2012) between an estimated and a true blur kernel."""

    def __str__(self):
        return "MNC"

    @staticmethod
    def calculate_mnc(estimated_kernel: torch.Tensor, true_kernel: torch.Tensor) -> torch.Tensor:
        assert estimated_kernel.shape[1] == 1 and true_kernel.shape[1] == 1, "kernels must have one channel"
        v = F.conv2d(estimated_kernel, true_kernel, padding="same")
        v = v / (torch.linalg.norm(estimated_kernel[0, 0], ord=2) * torch.linalg.norm(true_kernel[0, 0], ord=2))
        return v[0, 0].max()

    def preprocessing(self, img):
        return img.mean(dim=1, keepdim=True)

    def metric_fn(self, label, img):
        return self.calculate_mnc(img, label).item()


def run(input_dir, label_dir, exp_name, log_path="./result.log", gpu=0, metrics: Sequence[str] = ("FID", "LPIPS", "PSNR")):
    device = torch.device(f"cuda:{gpu}") if torch.cuda.is_available() else torch.device("cpu")
    logger = prepare_logger(log_path)
    table = {"FID": FID, "LPIPS": LPIPS, "PSNR": PSNR, "MSE": MSE, "MNC": MNC}
    logger.info(f"============= Metric Calculation for {exp_name} =============")
    return {name: table[name](input_dir, label_dir, logger, device).compute() for name in metrics}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--input_dir", type=str)
    parser.add_argument("--label_dir", type=str)
    parser.add_argument("--exp_name", type=str)
    parser.add_argument("--log_path", type=str, default="./result.log")
    parser.add_argument("--gpu", type=int, default=0)
    args = parser.parse_args()
    run(args.input_dir, args.label_dir, args.exp_name, args.log_path, args.gpu)


if __name__ == "__main__":
    main()
