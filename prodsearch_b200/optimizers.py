"""Optimizer wrapper with the reference's semantics (models/optimizers.py:111-245): Adam with
eps=1e-9 (:186), optional noam schedule (:214-219) and global-norm gradient clipping (:241-242).
Out of scope for new kernels ("next" row N2): it consumes the dense ``.grad``
tensors the gradient sinks attach, exactly like the reference's trainer (trainer.py:76-78)."""
import torch
from torch.nn.utils import clip_grad_norm_


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
            self.optimizer = torch.optim.Adam(self.params, lr=self.learning_rate, betas=self.betas, eps=1e-9,
                                              weight_decay=self.weight_decay)

    def step(self):
        self._step += 1
        if self.decay_method == "noam":
            self.learning_rate = self.original_lr * min(self._step ** (-0.5),
                                                        self._step * self.warmup_steps ** (-1.5))
            self.optimizer.param_groups[0]["lr"] = self.learning_rate
        if self.max_grad_norm:
            clip_grad_norm_(self.params, self.max_grad_norm)
        self.optimizer.step()


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
