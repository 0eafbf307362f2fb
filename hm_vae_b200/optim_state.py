"""``optimizer.pt`` in the reference's format: the ``state_dict()`` of ``torch.optim.Adam`` (trainer_motion_vae.py:29-31 builds
the optimiser, :112-113 / :121-126 save and load ``{'gen': gen_opt.state_dict()}``).

The fused optimisers keep their moments in flat arenas / per-tensor lists; these two pure functions convert between that and
torch's layout ``{'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [{lr, betas, eps, weight_decay, ...,
'params': [0..n-1]}]}`` so that a checkpoint written here loads into ``torch.optim.Adam`` of the reference and vice versa.
No CUDA needed (covered by the CPU tests).
"""
import torch


def _default_group(lr, betas, eps, weight_decay):
    """The hyper-parameter keys of this torch version's Adam param_group (amsgrad, foreach, capturable, ... vary by version)."""
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
    return {k: v for k, v in ref.param_groups[0].items() if k != "params"}


def to_torch_adam(step, lr, betas, eps, weight_decay, exp_avg, exp_avg_sq, live=None, initial_lr=None):
    """-> torch.optim.Adam.state_dict() layout.  ``live``: indices of parameters that have been stepped at least once (torch
    keeps no state entry for a parameter that never received a gradient); None = all."""
    n = len(exp_avg)
    idx = range(n) if live is None else sorted(live)
    state = {}
    for i in idx:
        state[i] = {"step": torch.tensor(float(step)), "exp_avg": exp_avg[i].detach().clone(),
                    "exp_avg_sq": exp_avg_sq[i].detach().clone()}
    group = _default_group(lr, betas, eps, weight_decay)
    if initial_lr is not None:
        group["initial_lr"] = initial_lr            # what torch.optim.lr_scheduler.StepLR adds to the group
    group["params"] = list(range(n))
    return {"state": state, "param_groups": [group]}


def from_torch_adam(sd, n):
    """torch layout (or this package's round-1 layout {step, lr, exp_avg: [...], exp_avg_sq: [...]}) ->
    (step, lr, exp_avg list with None for absent entries, exp_avg_sq list, live index set)."""
    if "state" in sd and "param_groups" in sd:
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != n:
            raise ValueError("optimizer state has %d parameters, the model has %d trainable ones" % (len(order), n))
        m, v, live, step = [None] * n, [None] * n, set(), 0
        for pos, key in enumerate(order):
            st = sd["state"].get(key)
            if not st:
                continue
            m[pos], v[pos] = st["exp_avg"], st["exp_avg_sq"]
            live.add(pos)
            step = max(step, int(float(st["step"])))
        return step, float(groups[0]["lr"]), m, v, live
    step = int(sd["step"])
    return step, sd.get("lr"), list(sd["exp_avg"]), list(sd["exp_avg_sq"]), set(range(n))
