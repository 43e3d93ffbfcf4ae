"""Epsilon thresholding of the anomaly scores on the device (reference eval_methods.py:186-236 `find_epsilon`, used by
`epsilon_eval` on A_Score_Global): the scores stay where the single-pass scorer left them."""
import torch

from ._lib import lib, check


def find_epsilon(scores, reg_level=1):
    """scores: 1-D CUDA float32 tensor.  Returns (epsilon, z, score) as Python floats; z = -1 when no candidate
    qualified (epsilon = max(scores), as the reference does)."""
    if not scores.is_cuda:
        raise ValueError("find_epsilon: CUDA tensor expected (host arrays: use the reference's eval_methods.find_epsilon)")
    s = scores.reshape(-1).contiguous().float()
    out = torch.empty(3, dtype=torch.float32, device=s.device)
    scratch = torch.empty(int(lib.mtadgat_find_epsilon_scratch_doubles(s.numel())), dtype=torch.float64, device=s.device)
    with torch.cuda.device(s.device):
        check(lib.mtadgat_find_epsilon(s.data_ptr(), s.numel(), int(reg_level), out.data_ptr(), scratch.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream))
    eps, z, sc = out.tolist()
    return eps, z, sc


def epsilon_predict(test_scores, train_scores, reg_level=1):
    """Point-wise predictions `score > epsilon` with epsilon from the training scores (epsilon_eval without the
    label-dependent point adjustment, eval_methods.py:165-168)."""
    eps, _, _ = find_epsilon(train_scores, reg_level)
    return test_scores > eps, eps
