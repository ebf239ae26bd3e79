"""chainer.optimizers.Adam as a torch optimizer (device-side, no host reads).

Update rule of chainer v6/v7 ``optimizers/adam.py`` (third-party, PARITY UNPINNED):
    m += (1 - beta1) (g - m);  v += (1 - beta2) (g^2 - v)
    alpha_t = alpha sqrt(1 - beta2^t) / (1 - beta1^t)
    p -= eta (alpha_t m / (sqrt(v) + eps) + weight_decay_rate p)
with t starting at 1.  torch.optim.Adam divides by sqrt(v) / sqrt(1 - beta2^t) + eps: not the
same numbers.  ``alpha`` is per parameter group, like chainer's per-parameter hyperparameters
(``link.translation.update_rule.hyperparam.alpha *= 0.1``)."""

import math

import torch


class ChainerAdam(torch.optim.Optimizer):
    def __init__(self, params, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8, eta=1.0,
                 weight_decay_rate=0.0):
        super().__init__(params, dict(alpha=alpha, beta1=beta1, beta2=beta2, eps=eps, eta=eta,
                                      weight_decay_rate=weight_decay_rate))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["beta1"], group["beta2"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["t"] = 0
                    st["m"] = torch.zeros_like(p)
                    st["v"] = torch.zeros_like(p)
                st["t"] += 1
                g, m, v = p.grad, st["m"], st["v"]
                m.add_(g - m, alpha=1.0 - b1)
                v.add_(g * g - v, alpha=1.0 - b2)
                alpha_t = group["alpha"] * math.sqrt(1.0 - math.pow(b2, st["t"])) / \
                    (1.0 - math.pow(b1, st["t"]))
                step = alpha_t * m / (v.sqrt() + group["eps"])
                if group["weight_decay_rate"]:
                    step = step + group["weight_decay_rate"] * p
                p.sub_(step, alpha=group["eta"])
        return loss
