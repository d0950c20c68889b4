"""Attentional FM oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED BY THE REFERENCE.

The reference's models/models_online_deep/afm_adam.py cannot run (a float is passed to .view at :67,69, undefined
self.verbose / evaluate / eval_metric at :121-123,167), so there is nothing to pin against.  This file
is the DEFINITION the CUDA AFMAdam is checked against instead: the model of the AFM paper (Xiao et al., IJCAI 2017, eq. 8)
with the reference's parameter set (afm_adam.py:34-41: first/second-order embeddings, bias, attention_linear = Linear(k, A),
H = randn(A), P = randn(k)), written in plain PyTorch (CPU, autograd), and the family's per-call fresh-state Adam step
 as the update rule:

    e_i   = x_i * V_i                                     (k-vector per field)
    z_ij  = e_i (.) e_j                      for i < j     (pairwise interaction layer)
    a'_ij = H . relu(W z_ij + c) ;  a_ij = softmax over the F(F-1)/2 pairs
    logit = bias + sum_i x_i w_i + sum_{i<j} a_ij * (P . z_ij)
"""
import numpy as np
import torch
import torch.nn.functional as F


class AFMTorch(torch.nn.Module):
    def __init__(self, feature_sizes, embedding_size=4, attention_size=4, b=0.99, n=0.003):
        super().__init__()
        self.feature_sizes, self.k, self.A = list(feature_sizes), embedding_size, attention_size
        R = int(sum(feature_sizes))
        self.offsets = torch.tensor(np.concatenate([[0], np.cumsum(feature_sizes)])[:-1], dtype=torch.long)
        self.w1 = torch.nn.Parameter(torch.zeros(R))
        self.V = torch.nn.Parameter(torch.zeros(R, embedding_size))
        self.bias = torch.nn.Parameter(torch.tensor(b))
        self.W = torch.nn.Parameter(torch.zeros(attention_size, embedding_size))
        self.c = torch.nn.Parameter(torch.zeros(attention_size))
        self.H = torch.nn.Parameter(torch.zeros(attention_size))
        self.P = torch.nn.Parameter(torch.zeros(embedding_size))
        self.n = torch.nn.Parameter(torch.tensor(n), requires_grad=False)
        Fd = len(feature_sizes)
        self.pi, self.pj = np.triu_indices(Fd, 1)

    def forward(self, Xi, Xv):
        ids = torch.as_tensor(np.asarray(Xi), dtype=torch.long).reshape(-1, len(self.feature_sizes)) + self.offsets
        xv = torch.as_tensor(np.asarray(Xv), dtype=torch.float32).reshape(ids.shape)
        first = (self.w1[ids] * xv).sum(1)
        e = self.V[ids] * xv.unsqueeze(-1)                      # [B,F,k]
        z = e[:, self.pi, :] * e[:, self.pj, :]                 # [B,P,k]
        att = F.relu(z @ self.W.t() + self.c) @ self.H          # [B,P]
        a = torch.softmax(att, dim=1)
        return self.bias + first + (a * (z @ self.P)).sum(1)

    def step(self, Xi, Xv, Y):
        """one batch: BCE-with-logits (mean), fresh-state Adam step on every parameter with a gradient; returns
        (loss, gradients dict) -- the gradients are what the CUDA kernels are compared on (the sign step itself is pinned
        on torch.optim.Adam in tests/test_oracle_math.py)."""
        opt = torch.optim.Adam([p for p in self.parameters() if p.requires_grad], lr=self.n)
        opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(self.forward(Xi, Xv), torch.as_tensor(np.asarray(Y), dtype=torch.float32))
        loss.backward()
        grads = {n_: p.grad.detach().clone().numpy() for n_, p in self.named_parameters() if p.grad is not None}
        opt.step()
        return float(loss.detach()), grads
