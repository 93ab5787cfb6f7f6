"""Algorithmic work of the planning hot path: the numerators of bench.py's roofline numbers.
FLOP = 2 * MAC over the Linear layers only (biases, activations, LayerNorm excluded)."""


def mlp_flops(L, A, M):
    """(F_dyn, F_rew, F_pi, F_Q) per row of TOLD's four 3-layer stacks (tdmpc.py:15-19, helper.py:169-176,197-201)."""
    f_dyn = 2 * ((L + A) * M + M * M + M * L)
    f_rew = 2 * ((L + A) * M + M * M + M)
    f_pi = 2 * (L * M + M * M + M * A)
    return f_dyn, f_rew, f_pi, f_rew


def gru_flops(L, A, M, Hd):
    """(F_step, F_rew) per row-step of the GRU latent dynamics (tdmpc_icem_similarity_drnn.py:56-59): gates + prior MLP, reward."""
    return 2 * 3 * Hd * (L + A + Hd) + 2 * (Hd * M + M * L), 2 * (Hd * M + M * M + M)


def flops_per_plan(D, A, L, M, E, H, rows_per_iter, P, Hd=0):
    """One plan(): sum_i T_i (H (F_dyn + F_rew) + F_pi + 2 F_Q) + P H (F_pi + F_dyn) + F_enc.  Hd > 0: GRU dynamics."""
    f_dyn, f_rew, f_pi, f_q = mlp_flops(L, A, M)
    if Hd:
        f_dyn, f_rew = gru_flops(L, A, M, Hd)
    tot = 2 * (D * E + E * L) + P * H * (f_pi + f_dyn)
    for T in rows_per_iter:
        tot += T * (H * (f_dyn + f_rew) + f_pi + 2 * f_q)
    return tot


def value_launch_flops(B, T, H, A, L, M, Hd=0):
    """One value-mode launch of the rollout kernel over B * T trajectories."""
    f_dyn, f_rew, f_pi, f_q = mlp_flops(L, A, M)
    if Hd:
        f_dyn, f_rew = gru_flops(L, A, M, Hd)
    return float(B) * T * (H * (f_dyn + f_rew) + f_pi + 2 * f_q)


def sample_steps_per_plan(H, rows_per_iter, P):
    """One sample-step = one `next` evaluation of one trajectory at one t."""
    return sum(T * H for T in rows_per_iter) + P * H
