"""FusedAdamL2: torch.optim.Adam semantics as configured at cmrl/models/dynamics/base_dynamics.py:58-63
(L2-coupled weight decay, bias correction, eps outside the sqrt) as ONE kernel launch over a mechanism's flat
parameter arena (`EnsembleMLP.flat_parameters`)."""
import torch

from . import ops


class FusedAdamL2(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.model = model
        params = list(model.parameters())
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = 0
        self._step_dev = None  # device copy of the step counter (CUDA-graph replayable steps)
        self._exp_avg = None
        self._exp_avg_sq = None

    def _arena(self):
        flat, flat_grad = self.model.flat_parameters()
        n = self.model._n_trainable
        if self._exp_avg is None or self._exp_avg.numel() != n or self._exp_avg.device != flat.device:
            self._exp_avg = torch.zeros(n, device=flat.device, dtype=torch.float32)
            self._exp_avg_sq = torch.zeros(n, device=flat.device, dtype=torch.float32)
        return flat[:n], flat_grad[:n]

    def zero_grad(self, set_to_none: bool = False):
        """Zeroes the flat gradient buffer; parameter .grad stay views of it (never set to None)."""
        _, flat_grad = self.model.flat_parameters()
        flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        p, grad = self._arena()
        if self._step_dev is not None:  # a graph replay may have advanced the device counter
            self._step = int(self._step_dev[0].item())
        self._step += 1
        ops.adam_l2_step(p, grad, self._exp_avg, self._exp_avg_sq, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                         g["weight_decay"], self._step, grad_scale)  # fmt: skip
        if self._step_dev is not None:
            # the device counter is never freed or replaced: captured step graphs hold its address
            self._step_dev[:1].fill_(self._step)
        self._mark_weights_changed()
        return loss

    def _mark_weights_changed(self):
        """The kernel writes the arena through raw pointers: torch's version counters do not move, so the model's own
        counter does (EnsembleMLP.weights_token; a replayed graph that contains the step calls this once per replay)."""
        holders = self.__dict__.get("_epoch_holders")
        if holders is None:  # the model and any wrapped mechanism (ForwardEulerTransition holds its one-step model)
            if not hasattr(self.model, "_weights_epoch"):
                self.model._weights_epoch = 0
            holders = self.__dict__["_epoch_holders"] = [m for m in self.model.modules() if hasattr(m, "_weights_epoch")]
        for m in holders:
            m._weights_epoch += 1

    def prepare_graphable(self):
        """Allocate the moment buffers and the device step counter (must happen outside a graph capture)."""
        p, _ = self._arena()
        if self._step_dev is None:  # [steps taken, arrival ticket of the kernel's last-block update]
            self._step_dev = torch.tensor([self._step, 0], dtype=torch.int32, device=p.device)

    @torch.no_grad()
    def step_graphable(self, grad_scale: float = 1.0):
        """Same update with the step counter on the device: safe to capture in a CUDA graph and replay."""
        g = self.param_groups[0]
        self.prepare_graphable()
        p, grad = self._arena()
        ops.adam_l2_step_dev(p, grad, self._exp_avg, self._exp_avg_sq, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                             g["weight_decay"], self._step_dev, grad_scale)  # fmt: skip
        self._mark_weights_changed()

    @torch.no_grad()
    def step_synced(self, peer, grad_scale: float = 1.0):
        """Data-parallel step: the gradient exchange over NVLink peer memory and this update in ONE launch
        (`cmrl_allreduce_grads_adam`); step counter on the device, so the call is graph-capturable."""
        g = self.param_groups[0]
        self.prepare_graphable()
        p, grad = self._arena()
        peer.allreduce_adam(p, grad, self._exp_avg, self._exp_avg_sq, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                            g["weight_decay"], self._step_dev, grad_scale)  # fmt: skip
        self._mark_weights_changed()

    @property
    def step_count(self):
        return self._step if self._step_dev is None else int(self._step_dev[0].item())

    def state_dict(self):
        return {
            "step": self.step_count,
            "exp_avg": None if self._exp_avg is None else self._exp_avg.clone(),
            "exp_avg_sq": None if self._exp_avg_sq is None else self._exp_avg_sq.clone(),
            "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
        }

    def load_state_dict(self, state):
        self._step = state["step"]
        if self._step_dev is not None:
            self._step_dev[:1].fill_(self._step)
        for name in ("exp_avg", "exp_avg_sq"):
            new, cur = state[name], getattr(self, "_" + name)
            if new is not None and cur is not None and cur.shape == new.shape and cur.device == new.device:
                cur.copy_(new)  # in place: captured step graphs hold the moment buffers' addresses
            else:
                setattr(self, "_" + name, None if new is None else new.clone())
        for g, s in zip(self.param_groups, state["param_groups"]):
            g.update(s)
