"""Minimal ``torch_ema.ExponentialMovingAverage`` work-alike for checkpoint compatibility.

The reference keeps an EMA of the model parameters, saves it under the checkpoint key ``"ema"``
and copies the shadow parameters into the live ones in ``eval()`` (reference universe.py:122-124,
841-869; model_loader.py:119-130).  Only the state layout (``shadow_params`` as a positional list
in ``model_parameters()`` order) and store / copy_to / restore are needed for inference.
"""
import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        parameters = list(parameters)
        self.shadow_params = [p.clone().detach() for p in parameters]
        self.collected_params = None

    def update(self, parameters):
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        with torch.no_grad():
            for s, p in zip(self.shadow_params, parameters):
                s.sub_((1.0 - decay) * (s - p))

    def copy_to(self, parameters):
        # in-place through the parameter itself (not ``p.data``): bumps ``p._version`` so that
        # ``engine.runtime.weights_version`` notices the swap and re-packs the device weights
        with torch.no_grad():
            for s, p in zip(self.shadow_params, parameters):
                p.copy_(s)

    def store(self, parameters):
        self.collected_params = [p.detach().clone() for p in parameters]

    def restore(self, parameters):
        if self.collected_params is None:
            raise RuntimeError("no stored parameters to restore")
        with torch.no_grad():
            for c, p in zip(self.collected_params, parameters):
                p.copy_(c)
        self.collected_params = None

    def to(self, device=None, dtype=None):
        self.shadow_params = [p.to(device=device, dtype=dtype) if p.is_floating_point()
                              else p.to(device=device) for p in self.shadow_params]
        if self.collected_params is not None:
            self.collected_params = [p.to(device=device) for p in self.collected_params]

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates,
                "shadow_params": self.shadow_params, "collected_params": self.collected_params}

    def load_state_dict(self, state_dict):
        self.decay = state_dict["decay"]
        self.num_updates = state_dict.get("num_updates")
        shadow = state_dict["shadow_params"]
        if len(shadow) != len(self.shadow_params):
            raise ValueError("EMA state has a different number of parameters than the model")
        self.shadow_params = [s.to(p.device, p.dtype).clone()
                              for s, p in zip(shadow, self.shadow_params)]
        coll = state_dict.get("collected_params")
        self.collected_params = None if coll is None else [c.clone() for c in coll]
