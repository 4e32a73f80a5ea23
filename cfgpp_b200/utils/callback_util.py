"""Per-step callback protocol of the reference (utils/callback_util.py:8-37, 67-75): a callback is called as
`callback_fn(step, t, {'z0t', 'zt', 'decode'}) -> dict` at the end of a step and may replace `z0t` / `zt`.
Installing one makes the solvers use the un-fused seam so both tensors are materialised every step. The two
image-dumping callbacks of the reference (draw_tweedie / draw_noisy, :39-65) decode the Tweedie estimate / the noisy
latent of a step and write it under <workdir>/record/."""
from pathlib import Path

import torch

__CALLBACK__ = {}


def register_callback(name):
    def wrapper(cls):
        if __CALLBACK__.get(name) is not None:
            raise NameError(f"Callback {name} is already registered")
        __CALLBACK__[name] = cls
        return cls
    return wrapper


def get_callback(name, **kwargs):
    if __CALLBACK__.get(name) is None:
        raise NameError(f"Callback {name} is not registered")
    return __CALLBACK__[name](**kwargs)


class DiffusionCallback:
    def __init__(self, frequency: int, workdir: Path):
        assert frequency > 0, "Frequency must be a positive float"
        self.frequency = frequency
        self.workdir = workdir

    def __call__(self, step, t, callback_kwargs):
        if (step + 1) % self.frequency == 0 or step == 0:
            return self.callback(step, t, callback_kwargs)
        return callback_kwargs

    def callback(self, step, t, callback_kwargs):
        raise NotImplementedError


@register_callback("record")
class RecordCallback(DiffusionCallback):
    """Keeps (step, t, z0t, zt) on the host — the in-memory analogue of the reference's PNG-dumping callbacks."""
    def __init__(self, frequency: int = 1, workdir: Path = None):
        super().__init__(frequency, workdir)
        self.records = []

    def callback(self, step, t, callback_kwargs):
        self.records.append((step, int(t), callback_kwargs["z0t"].detach().cpu(), callback_kwargs["zt"].detach().cpu()))
        return callback_kwargs


class _DrawCallback(DiffusionCallback):
    """Decode one of the step's latents and dump it (PNG through torchvision when importable, the tensor otherwise)."""
    key, folder, prefix = "z0t", "tweedie", "x0"

    def __init__(self, frequency: int, workdir: Path):
        super().__init__(frequency, workdir)
        self.outdir = Path(workdir).joinpath("record", self.folder)
        self.outdir.mkdir(parents=True, exist_ok=True)

    @torch.no_grad()
    def callback(self, step, t, callback_kwargs):
        img = callback_kwargs["decode"](callback_kwargs[self.key])
        img = (img / 2 + 0.5).clamp(0, 1).cpu()
        stem = self.outdir.joinpath(f"{self.prefix}_{int(t)}")
        try:
            from torchvision.utils import save_image
            save_image(img, stem.with_suffix(".png"))
        except Exception:  # torchvision is optional in this image
            torch.save(img, stem.with_suffix(".pt"))
        return callback_kwargs


@register_callback("draw_tweedie")
class DrawTweedieCallback(_DrawCallback):
    key, folder, prefix = "z0t", "tweedie", "x0"


@register_callback("draw_noisy")
class DrawNoisyCallback(_DrawCallback):
    key, folder, prefix = "zt", "noisy", "xt"


class ComposeCallback(DiffusionCallback):
    def __init__(self, workdir, callbacks, frequency: int = 5):
        super().__init__(frequency, workdir)
        self.callbacks = [get_callback(name, workdir=workdir, frequency=frequency) for name in callbacks]

    def __call__(self, step, t, callback_kwargs):
        for callback in self.callbacks:
            callback_kwargs = callback(step, t, callback_kwargs)
        return callback_kwargs
