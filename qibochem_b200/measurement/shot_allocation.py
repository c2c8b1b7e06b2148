"""
Host-side shot budgeting (reference: ``src/qibochem/measurement/shot_allocation.py``).  Pure integer policy; the
``shot_allocation=`` argument of ``expectation_from_samples`` is what reaches the device path.
"""

from __future__ import annotations

import numpy as np


def coefficients_sum(expression) -> float:
    """Sum of |coefficient| over the non-constant terms of a group expression."""
    return float(sum(abs(c) for k, c in expression.terms.items() if k != (0, 0)))


def allocate_shots(grouped_terms, n_shots: int, method: str | None = None, max_shots_per_term: int | None = None) -> list[int]:
    """Split ``n_shots`` over the term groups (shot_allocation.py:14-99).

    ``method`` "c"/"coefficients" (default): proportional to (sum |c|)^(2/3) of each group (arXiv:2307.06504), capped
    by ``max_shots_per_term``; "u"/"uniform": evenly, leftovers placed at random.
    """
    method = "c" if method is None else method
    weights = np.array([coefficients_sum(expr) ** (2.0 / 3.0) for expr, *_ in grouped_terms], dtype=float)
    if max_shots_per_term is None:
        max_shots_per_term = int(np.ceil(n_shots * (weights.max() / weights.sum())))
    cap = min(n_shots, max_shots_per_term)
    n_groups = len(grouped_terms)
    alloc = np.zeros(n_groups, dtype=int)
    while True:
        left = n_shots - int(alloc.sum())
        if not left:
            break
        if alloc.min() == cap:  # every group is full but shots remain: double the cap and start again
            cap = min(2 * cap, n_shots)
            alloc[:] = 0
            continue
        if method in ("c", "coefficients"):
            w = np.where(alloc < cap, weights, 0.0)
            w = w / w.sum()
            step = (left * w).astype(int)
            w = w * (step > 0)
            if step.any():
                w = w / w.sum()
                step = (left * w).astype(int)
            else:  # fewer shots left than groups: hand them out one by one
                step = np.array(allocate_shots(grouped_terms, left, method="u", max_shots_per_term=left))
        elif method in ("u", "uniform"):
            step = np.full(n_groups, left // n_groups)
            if not step.any():
                step = np.zeros(n_groups)
                step[:left] = 1
                np.random.shuffle(step)
        else:
            raise NameError("Unknown method!")
        alloc = np.clip(alloc + step.astype(int), 0, cap)
    return alloc.tolist()


def allocate_shots_by_variance(total_shots: int, n_trial_shots: int, variance_values, method: str = "vmsa") -> list[int]:
    """Distribute the shots left after the trial runs in proportion to the groups' sample standard deviations
    (VMSA), optionally scaled down by eta < 1 (VPSR)  (shot_allocation.py:102-125)."""
    assert method in ("vmsa", "vpsr"), f"Unknown shot assignment method ({method}) called"
    n_groups = len(variance_values)
    left = total_shots - n_groups * n_trial_shots
    sigma = [v**0.5 for v in variance_values]
    eta = 1 if method == "vmsa" else sum(sigma) ** 2 / (n_groups * sum(variance_values))
    out = [int(eta * s * left / sum(sigma)) for s in sigma]
    if method == "vmsa":
        out[-1] += left - sum(out)
    return out
