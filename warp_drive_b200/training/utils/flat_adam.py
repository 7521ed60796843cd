"""Adam + gradient-norm clipping over ONE flat float32 arena per policy.

Every trained tensor of a policy becomes a VIEW into `params`, every `.grad` a view into
`grads`: the NCCL gradient all-reduce runs on the flat gradient buffer in place (no pack /
unpack copies), clipping is one reduction kernel, Adam one elementwise kernel
(`wdb_grad_sumsq`, `wdb_adam_step`, csrc/wdb_update.cu), and nothing synchronises with the
host.  Same update rule as torch.optim.Adam(lr) + torch.nn.utils.clip_grad_norm_, which the
reference calls per tensor (trainer_a2c.py:300-339).
"""
import torch

from warp_drive_b200 import lib as _lib


def _slot(k):
    """Arena elements reserved for a tensor of k elements: every tensor starts on a 16-byte
    boundary (cuBLASLt's fused epilogues and 16-byte loads need aligned weight / bias
    pointers); the padding holds zeros, receives zero gradients and never moves."""
    return (k + 3) // 4 * 4


class FlatAdam:
    @staticmethod
    def numel(parameters):
        return sum(_slot(p.numel()) for p in parameters if p.requires_grad)

    def __init__(self, parameters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, arena=None):
        """arena: optional (params, grads) flat float32 tensors of exactly numel(parameters)
        elements (numel() counts every tensor rounded up to 4 elements) (slices of a larger arena shared by several policies, so that ONE
        all-reduce covers all of them)."""
        self.param_list = [p for p in parameters if p.requires_grad]
        assert self.param_list, "no trainable parameters"
        dev = self.param_list[0].device
        assert dev.type == "cuda", "FlatAdam runs on the GPU (no CPU fallback)"
        n = sum(_slot(p.numel()) for p in self.param_list)
        self.n = n
        if arena is not None:
            self.params, self.grads = arena
            assert self.params.numel() == n and self.grads.numel() == n
            assert self.params.is_contiguous() and self.grads.is_contiguous()
            self.grads.zero_()
        else:
            self.params = torch.zeros(n, dtype=torch.float32, device=dev)
            self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        off = 0
        for p in self.param_list:
            k = p.numel()
            self.params[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.params[off:off + k].view_as(p)
            p.grad = self.grads[off:off + k].view_as(p)
            off += _slot(k)
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.step_count = 0
        # torch.optim-style handle so that schedulers / callers can set group["lr"]
        self.param_groups = [{"lr": self.lr, "params": self.param_list}]

    def zero_grad(self):
        self.grads.zero_()
        # autograd accumulates in place into the existing .grad views; re-attach any that a
        # caller replaced (e.g. zero_grad(set_to_none=True) elsewhere)
        off = 0
        for p in self.param_list:
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grads[off:off + k].data_ptr():
                p.grad = self.grads[off:off + k].view_as(p)
            off += _slot(k)

    def grad_norm(self):
        """Device tensor holding the 2-norm of the (unclipped) flat gradient."""
        self.sumsq.zero_()
        _lib.check(_lib.load().wdb_grad_sumsq(_lib.stream_ptr(), _lib.ptr(self.grads), self.n,
                                              _lib.ptr(self.sumsq)), "grad_sumsq")
        return self.sumsq.sqrt()

    def step(self, max_grad_norm=None):
        """One Adam step; `max_grad_norm` clips the global 2-norm of this arena's gradient
        first (clip_grad_norm_ semantics)."""
        self.step_count += 1
        lr = float(self.param_groups[0]["lr"])
        clip = float(max_grad_norm) if max_grad_norm else 0.0
        if clip > 0.0:
            self.sumsq.zero_()
            _lib.check(_lib.load().wdb_grad_sumsq(_lib.stream_ptr(), _lib.ptr(self.grads),
                                                  self.n, _lib.ptr(self.sumsq)), "grad_sumsq")
        _lib.check(_lib.load().wdb_adam_step(
            _lib.stream_ptr(), _lib.ptr(self.params), _lib.ptr(self.grads),
            _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), self.n, lr, self.betas[0],
            self.betas[1], self.eps, self.step_count, clip,
            _lib.ptr(self.sumsq) if clip > 0.0 else None), "adam_step")

    # checkpoint support (same keys torch.optim uses are not needed by the reference, which
    # only saves model state dicts)
    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(),
                "exp_avg_sq": self.exp_avg_sq.clone(), "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, state):
        self.step_count = int(state["step"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        self.param_groups[0]["lr"] = state.get("lr", self.param_groups[0]["lr"])
