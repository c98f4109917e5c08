"""Step program of the `--method linear` model (reference: linear_model.py:142-184 train / test / score loops,
torch.nn.MSELoss :236, torch.optim.Adam :242): y = x W^T + b over flattened images, MSE, Adam with coupled L2.

Composed from existing entry points of libcae_b200, launched eagerly per batch (six launches per optimiser step):
    cae_gemm          yhat = x W^T + b
    cae_mse           loss = mean((yhat - y)^2)
    cae_ew_epilogue   dz = 2 (yhat - y) / count      (on-load transform k0*t0 + k1*t1 of the two tensors)
    cae_gemm          dW = dz^T x, db = row sums of dz^T
    cae_adam          both parameters through the flat arena
    cae_step_advance
ops.gemm sends both contractions to the tensor cores (tc_dense.cu: split -> tcgen05 3xTF32 -> epilogue) when the batch is
>= 128; below that the layer is bound by one read of its weight (67 MB at 256 -> 65 536) and stays on the fp32 tile kernel."""

from __future__ import annotations

import torch

from . import ops
from .convae import DataBinding, _align
from .._lib import require_cuda


class LinearEngine:

    def __init__(self, module, lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8, device="cuda", grad_hook=None,
                 count_scale=1.0):
        """grad_hook / count_scale: data parallelism (engine/dp.py) - every rank runs its share of each batch, the loss and
        its gradient are scaled by count_scale = n_local / n_global, grad_hook SUM-all-reduces the flat gradient arena
        before the optimiser step (so the summed gradients and the summed losses are those of the global batch)"""
        require_cuda()
        self.grad_hook, self.count_scale = grad_hook, float(count_scale)
        self.device = torch.device(device)
        self.module = module.to(self.device)
        self.lin = module.linear[1]
        self.out_shape = tuple(module.output_shape)
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        params = [self.lin.weight, self.lin.bias]
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += _align(p.numel())
        z = lambda: torch.zeros(total, dtype=torch.float32, device=self.device)
        self.arena, self.grads, self.adam_m, self.adam_v = z(), z(), z(), z()
        self._g = {}
        with torch.no_grad():
            for p, off in zip(params, offs):
                n = p.numel()
                self.arena[off:off + n].copy_(p.detach().reshape(-1).to(self.device, torch.float32))
                p.data = self.arena[off:off + n].view(p.shape)
                self._g[id(p)] = self.grads[off:off + n].view(p.shape)
                p.grad = self._g[id(p)]
        self.n_params = total
        self.step_count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._partials = torch.zeros(ops.partials_len(1), dtype=torch.float64, device=self.device)
        self._bufs = {}

    def bind(self, X, Y, batch_size):
        X = X.to(self.device, torch.float32)
        Y = Y.to(self.device, torch.float32) if Y is not None else None
        return DataBinding(X, Y, batch_size)

    def _buffers(self, B):
        if B not in self._bufs:
            C = self.out_shape[0]
            f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
            self._bufs[B] = {"yhat": f(B, *self.out_shape), "dz": f(B, *self.out_shape), "kp": f(C), "kn": f(C)}
        return self._bufs[B]

    def _forward(self, x, yhat):
        N, K, O = x.shape[0], self.lin.in_features, self.lin.out_features
        ops.gemm(N, O, K, x, K, 1, self.lin.weight, 1, K, yhat, O, 1, bias=self.lin.bias)

    def _batches(self, data):
        B = data.batch_size
        for i in range(data.n_batches):
            lo, hi = i * B, min(data.n, (i + 1) * B)
            yield i, data.X[lo:hi].reshape(hi - lo, -1), (data.Y[lo:hi] if data.Y is not None else None)

    def train_epoch(self, data):
        K, O = self.lin.in_features, self.lin.out_features
        G = self._g
        for i, x, y in self._batches(data):
            N = x.shape[0]
            b = self._buffers(N)
            yhat, dz = b["yhat"][:N], b["dz"][:N]
            self._forward(x, yhat)
            ops.mse(yhat, y, y.numel(), self._partials, self._ticket, data.losses[i:i + 1])
            if self.count_scale != 1.0:
                data.losses[i:i + 1].mul_(self.count_scale)
            b["kp"].fill_(2.0 * self.count_scale / y.numel())
            b["kn"].fill_(-2.0 * self.count_scale / y.numel())
            ops.ew_epilogue(ops.make_src(yhat, t1=y.contiguous(), k0=b["kp"], k1=b["kn"]), ops.view4(dz),
                            ops.make_epilogue(ops.EPI_PLAIN))
            # dW[o][k] = sum_n dz[n][o] x[n][k] ; db[o] = sum_n dz[n][o]
            ops.gemm(O, K, N, dz, 1, O, x, K, 1, G[id(self.lin.weight)], K, 1, rowsum_A=G[id(self.lin.bias)])
            if self.grad_hook is not None:
                self.grad_hook(self.grads)
            ops.adam(self.arena, self.grads, self.adam_m, self.adam_v, self.n_params, self.lr, self.betas[0],
                     self.betas[1], self.eps, self.weight_decay, False, 1.0, self.step_count)
            ops.step_advance(self.step_count, None, 1)
        return data.losses

    def test_epoch(self, data):
        for i, x, y in self._batches(data):
            yhat = self._buffers(x.shape[0])["yhat"][:x.shape[0]]
            self._forward(x, yhat)
            ops.mse(yhat, y, y.numel(), self._partials, self._ticket, data.losses[i:i + 1])
            if self.count_scale != 1.0:
                data.losses[i:i + 1].mul_(self.count_scale)
        return data.losses

    def score_batches(self, data, sink):
        for i, x, _ in self._batches(data):
            yhat = self._buffers(x.shape[0])["yhat"][:x.shape[0]]
            self._forward(x, yhat)
            sink(i, yhat)
