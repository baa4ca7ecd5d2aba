"""Optimizer wrapper with the reference's semantics (models/optimizers.py:111-245): Adam with
eps=1e-9 (:186), optional noam schedule (:214-219) and global-norm gradient clipping (:241-242).
Row N2 of SURVEY.md 8(f): the clip + Adam update of every parameter tensor runs as ONE fused multi-tensor
call (psb_adam_step) on the dense ``.grad`` tensors the gradient sinks attach, with the step counter and the
global norm in device memory (CUDA-graph replayable).  Dense semantics, identical to the reference's."""
import torch
from torch.nn.utils import clip_grad_norm_

from . import _lib


class FusedAdam(object):
    """torch.optim.Adam(eps=1e-9) + clip_grad_norm_ in three sm_100a launches (psb_adam_step).  Keeps the
    ``param_groups`` / ``state_dict`` / ``load_state_dict`` surface of torch.optim.Adam that the reference's
    checkpoints use (models/ps_model.py:27-32, trainer.py:234-244)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-9, weight_decay=0.0):
        self.params = list(params)
        self.param_groups = [dict(params=self.params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                  amsgrad=False)]
        self.state = {}
        self._step_dev = None
        self._sqnorm = None
        self._ws = None
        self._norm_given = 0

    def _ensure(self, dev):
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            self._sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)

    def _state_of(self, p):
        st = self.state.get(p)
        if st is None:
            st = self.state[p] = dict(exp_avg=torch.zeros_like(p, memory_format=torch.contiguous_format),
                                      exp_avg_sq=torch.zeros_like(p, memory_format=torch.contiguous_format))
        return st

    def set_global_sqnorm(self, sq):
        """Row-sharded training: the caller supplies the GLOBAL squared gradient norm (device tensor [1]; this
        rank's parameters are only a shard) for the next ``step``.  Stream-ordered copy, CUDA-graph safe."""
        self._ensure(sq.device)
        self._sqnorm.copy_(sq.reshape(1))
        self._norm_given = 1

    def global_norm_slots(self, device):
        """(sqnorm, step) device tensors a caller fills itself (psb_peer_sum_sqnorm) for the next ``step``."""
        self._ensure(device)
        self._norm_given = 2
        return self._sqnorm, self._step_dev

    @property
    def total_norm(self):
        """Global gradient norm of the last step (device tensor; reading it synchronises)."""
        return self._sqnorm.sqrt()

    def step(self, max_grad_norm=0.0, noam=False, warmup_steps=4000.0):
        g = self.param_groups[0]
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        if len(live) > _lib.ADAM_MAX_TENSORS:
            raise RuntimeError("FusedAdam: more than %d parameter tensors" % _lib.ADAM_MAX_TENSORS)
        self._ensure(live[0].device)
        arr = (_lib.AdamTensor * len(live))()
        for i, p in enumerate(live):
            st = self._state_of(p)
            grad = p.grad
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and grad.is_contiguous()
                    and grad.dtype == torch.float32 and not grad.is_sparse):
                raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters and dense gradients")
            arr[i] = _lib.AdamTensor(p.data_ptr(), grad.data_ptr(), st["exp_avg"].data_ptr(),
                                     st["exp_avg_sq"].data_ptr(), p.numel())
        lib = _lib.load()
        wb = int(lib.psb_adam_workspace_bytes(arr, len(live)))
        if self._ws is None or self._ws.numel() < wb:
            self._ws = torch.empty(wb, dtype=torch.uint8, device=live[0].device)
        b1, b2 = g["betas"]
        from . import ops
        ev = None
        if ops.PROFILE is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.check(lib.psb_adam_step(arr, len(live), float(g["lr"]), float(b1), float(b2), float(g["eps"]),
                                     float(g["weight_decay"]), float(max_grad_norm or 0.0), 1 if noam else 0,
                                     float(warmup_steps), int(self._norm_given), self._step_dev.data_ptr(),
                                     self._sqnorm.data_ptr(),
                                     self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()), "psb_adam_step")
        self._norm_given = 0
        _lib.note_param_write()        # raw-pointer update: tensor._version does not move
        if ev is not None:
            ev[1].record()
            ops.PROFILE.setdefault("adam_step", []).append(ev)

    # ---- torch.optim.Adam-compatible checkpoint format ----------------------------------------
    def state_dict(self):
        step = int(self._step_dev.item()) if self._step_dev is not None else 0
        state = {}
        for i, p in enumerate(self.params):
            st = self.state.get(p)
            if st is not None:
                state[i] = dict(step=torch.tensor(float(step)), exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"])
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g["params"] = list(range(len(self.params)))
        return dict(state=state, param_groups=[g])

    def load_state_dict(self, sd):
        step = 0
        for i, st in sd["state"].items():
            p = self.params[int(i)]
            mine = self._state_of(p)
            mine["exp_avg"].copy_(st["exp_avg"])
            mine["exp_avg_sq"].copy_(st["exp_avg_sq"])
            step = max(step, int(float(st["step"])))
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
        if self.params:
            self._ensure(self.params[0].device)
            self._step_dev.fill_(step)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None


class Optimizer(object):
    def __init__(self, method, learning_rate, max_grad_norm, beta1=0.9, beta2=0.999, decay_method=None,
                 warmup_steps=4000, weight_decay=0.0):
        if method not in ("adam", "sgd"):
            raise RuntimeError("Invalid optim method: " + method)
        self.method = method
        self.learning_rate = learning_rate
        self.original_lr = learning_rate
        self.max_grad_norm = max_grad_norm
        self.betas = [beta1, beta2]
        self.decay_method = decay_method
        self.warmup_steps = warmup_steps
        self.weight_decay = weight_decay
        self._step = 0

    def set_parameters(self, params):
        self.params = [p for _, p in params if p.requires_grad]
        if self.method == "sgd":
            self.optimizer = torch.optim.SGD(self.params, lr=self.learning_rate, weight_decay=self.weight_decay)
        else:
            self.optimizer = FusedAdam(self.params, lr=self.learning_rate, betas=self.betas, eps=1e-9,
                                       weight_decay=self.weight_decay)

    def step(self):
        self._step += 1
        noam = self.decay_method == "noam"
        if noam:      # host mirror of the schedule the kernel evaluates from its device step counter
            self.learning_rate = self.original_lr * min(self._step ** (-0.5),
                                                        self._step * self.warmup_steps ** (-1.5))
        if self.method == "sgd":
            if noam:
                self.optimizer.param_groups[0]["lr"] = self.learning_rate
            if self.max_grad_norm:
                clip_grad_norm_(self.params, self.max_grad_norm)
            self.optimizer.step()
            return
        self.optimizer.param_groups[0]["lr"] = self.original_lr if noam else self.learning_rate
        self.optimizer.step(self.max_grad_norm, noam, self.warmup_steps)


def build_optim(args, model, checkpoint=None):
    """models/ps_model.py:20-51 (fresh optimizer; checkpoint restore reuses the saved object)."""
    if getattr(args, "train_from", "") and checkpoint is not None:
        optim = checkpoint["optim"]
        saved = optim.optimizer.state_dict()
        optim.set_parameters(list(model.named_parameters()))
        optim.optimizer.load_state_dict(saved)
        return optim
    optim = Optimizer(args.optim, args.lr, args.max_grad_norm, beta1=args.beta1, beta2=args.beta2,
                      decay_method=args.decay_method, warmup_steps=args.warmup_steps,
                      weight_decay=args.l2_lambda)
    optim.set_parameters(list(model.named_parameters()))
    return optim
