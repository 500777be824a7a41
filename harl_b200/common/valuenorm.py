"""ValueNorm with device-resident state (reference: harl/common/valuenorm.py:7-92).

The three running statistics live in one CUDA float[3] tensor that the GAE and value-loss
kernels read directly; ``update / normalize / denormalize`` keep the reference's signatures
(``denormalize`` returns NumPy, ``normalize`` a torch tensor) for callers outside the hot path.
"""
import numpy as np
import torch

from .. import _lib as L


class ValueNorm:
    def __init__(self, input_shape=1, norm_axes=1, beta=0.99999, per_element_update=False, epsilon=1e-5,
                 device=torch.device("cpu")):
        if input_shape != 1 or per_element_update or norm_axes != 1:
            raise NotImplementedError("only the scalar ValueNorm(1) the on-policy runner builds is supported")
        self.input_shape, self.norm_axes, self.beta, self.epsilon = input_shape, norm_axes, beta, epsilon
        self.device = torch.device(device)
        self.state = torch.zeros(3, dtype=torch.float32, device=self.device)  # mean, mean_sq, debias
        self._m3 = torch.zeros(3, dtype=torch.float64, device=self.device)

    # reference attribute names (valuenorm.py:28-36) as views of the state vector
    running_mean = property(lambda self: self.state[0:1])
    running_mean_sq = property(lambda self: self.state[1:2])
    debiasing_term = property(lambda self: self.state[2])

    def _dev(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        return x.to(device=self.device, dtype=torch.float32).contiguous()

    def running_mean_var(self):
        d = self.state[2].clamp(min=self.epsilon)
        mean = self.state[0] / d
        var = (self.state[1] / d - mean**2).clamp(min=1e-2)
        return mean.reshape(1), var.reshape(1)

    def update_from_moments(self, m3):
        """m3: device double[3] = (sum, sum of squares, count) of the batch (already reduced over ranks)."""
        L.call("hb_valuenorm_update", L.ptr(self.state), L.ptr(m3), float(self.beta), L.stream_ptr())

    @torch.no_grad()
    def update(self, input_vector):
        x = self._dev(input_vector)
        self._m3.zero_()
        L.call("hb_masked_moments", L.ptr(x), None, x.numel(), L.ptr(self._m3), L.stream_ptr())
        self.update_from_moments(self._m3)

    def _apply(self, x, denorm):
        x = self._dev(x)
        y = torch.empty_like(x)
        L.call("hb_valuenorm_apply", L.ptr(self.state), L.ptr(x), L.ptr(y), x.numel(), int(denorm), L.stream_ptr())
        return y

    def normalize(self, input_vector):
        return self._apply(input_vector, False)

    def denormalize(self, input_vector):
        return self._apply(input_vector, True).cpu().numpy()

    def state_dict(self):
        return {"running_mean": self.state[0:1].clone(), "running_mean_sq": self.state[1:2].clone(),
                "debiasing_term": self.state[2].clone()}

    def load_state_dict(self, sd):
        if not sd:  # the reference saves an empty dict when trained on CUDA
            return
        with torch.no_grad():
            self.state[0] = torch.as_tensor(sd["running_mean"]).reshape(-1)[0]
            self.state[1] = torch.as_tensor(sd["running_mean_sq"]).reshape(-1)[0]
            self.state[2] = torch.as_tensor(sd["debiasing_term"]).reshape(-1)[0]
