"""CPU check of the identity the CUDA rollout path relies on (rollout.cu / rollout_tc.cu).

The oracle (oracle/clip_standin.py:vision_attention_rollout) forms R = prod_l rownorm(0.5 * mean_h P_l + 0.5 * I) with N x N
matrix products and returns R[:, 0, 1:].  The kernels never build a map: they propagate the CLS row from the last layer to the
first, r <- 0.5 r + (0.5 / H) sum_h r^T P_{l,h}, recomputing P from q, k and the per-row statistics
lse = log2 sum_j 2^(c q.k_j), with r_i folded into the exponent.  Both formulations are restated here in plain torch (fp64)
and compared, including the exact zeros of the first step (r = e_0) and rows whose statistics are "unwritten" (NaN)."""
import math

import torch


def _oracle_rollout(p_layers):
    """p_layers: list of [B, H, N, N] probabilities, first layer first (same loop as the oracle's)."""
    n = p_layers[0].shape[-1]
    eye = torch.eye(n, dtype=p_layers[0].dtype)
    R = eye.expand(p_layers[0].shape[0], n, n).clone()
    for P in p_layers:
        A = 0.5 * P.mean(dim=1) + 0.5 * eye
        A = A / A.sum(dim=-1, keepdim=True)
        R = A @ R
    return R[:, 0, 1:]


def _recompute_rollout(q_layers, k_layers, dead_rows_last=0):
    """The kernels' formulation: log2-domain statistics, r folded into the exponent, last layer first."""
    c = math.log2(math.e) / 8.0
    L = len(q_layers)
    B, H, N, _ = q_layers[0].shape
    r = torch.zeros(B, N, dtype=q_layers[0].dtype)
    r[:, 0] = 1.0
    for l in range(L - 1, -1, -1):
        x = c * (q_layers[l] @ k_layers[l].transpose(-1, -2))                 # [B,H,N,N] log2-domain scores
        lse = torch.logsumexp(x * math.log(2.0), dim=-1) / math.log(2.0)     # log2 sum_j 2^x
        if l == L - 1 and dead_rows_last:
            lse[:, :, dead_rows_last:] = float("nan")                        # statistics of dead query rows are never written
        lr = torch.log2(r)                                                    # -inf where r == 0
        lw = torch.where(torch.isinf(lr)[:, None, :].expand_as(lse), torch.full_like(lse, float("inf")), lse - lr[:, None, :])
        contrib = torch.exp2(x - lw[..., None]).sum(dim=2)                   # sum over queries i -> [B,H,N(keys)]
        r = 0.5 * r + (0.5 / H) * contrib.sum(dim=1)
    return r[:, 1:]


def test_row_vector_recompute_equals_the_oracle_matrix_product():
    torch.manual_seed(0)
    B, H, N, L = 2, 3, 21, 5
    qs = [torch.randn(B, H, N, 64, dtype=torch.float64) for _ in range(L)]
    ks = [torch.randn(B, H, N, 64, dtype=torch.float64) for _ in range(L)]
    ps = [torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) for q, k in zip(qs, ks)]
    ref = _oracle_rollout(ps)
    out = _recompute_rollout(qs, ks)
    assert torch.allclose(out, ref, rtol=1e-10, atol=0)
    assert torch.allclose(out.sum(-1), ref.sum(-1))
    # last layer with dead-row elimination: only the CLS row's statistics exist, the rest must never be read
    out_dead = _recompute_rollout(qs, ks, dead_rows_last=1)
    assert torch.isfinite(out_dead).all() and torch.allclose(out_dead, ref, rtol=1e-10, atol=0)


def test_rollout_rows_sum_to_one_without_the_oracles_row_normalisation():
    """rows of mean_h P sum to 1, so rownorm(0.5 A + 0.5 I) is the identity operation the kernels omit."""
    torch.manual_seed(1)
    P = torch.softmax(torch.randn(2, 4, 9, 9, dtype=torch.float64), dim=-1)
    A = 0.5 * P.mean(1) + 0.5 * torch.eye(9, dtype=torch.float64)
    assert torch.allclose(A.sum(-1), torch.ones(2, 9, dtype=torch.float64), atol=1e-14)
