"""Algorithmic work of the path, frozen for the roofline report.

Counting rule: add/mul/div = 1 flop, fma = 2, transcendental (exp/log/sqrt) = 1 "T-op" tallied
separately.  Per (profile, level, frequency) and model with N_h H2O and N_o O2 lines:
  per-line per-frequency work  H2O c_h = 20 flop, O2 c_o = 22 flop; continua + N2 + non-resonant c_c = 60
  W_abs = P*L*F*(N_h*20 + N_o*22 + 60)            (= 1458*P*L*F for 16 + 49 lines)
  W_pre = P*L*(N_h*30 + N_o*14) flop  (+ P*L*(N_h*5 + N_o*1 + 4) T-ops)   frequency-independent
  W_rte = P*F*A*L*30 flop             (+ P*F*A*L*2 T-ops)
Algorithmic bytes: in 4*8*P*L (+16*P*L with cloud), out 6*8*P*F*A (+8*P*F*L with per-layer tau).
These are ALGORITHMIC counts (what the math needs), not executed-instruction counts: fp64 division,
exp and log expand to many DFMA-pipe instructions on the GPU, so "achieved flop/s" understates pipe
occupancy -- the ncu FP64-pipe utilisation in profiles/ is the measure north_star asks for.
"""

C_H2O, C_O2, C_CONT = 20, 22, 60


def work(P, L, F, A, n_h2o=16, n_o2=49, cloudy=False, tau_layers=False):
    w_abs = P * L * F * (n_h2o * C_H2O + n_o2 * C_O2 + C_CONT)
    w_pre = P * L * (n_h2o * 30 + n_o2 * 14)
    t_pre = P * L * (n_h2o * 5 + n_o2 * 1 + 4)
    w_rte = P * F * A * L * 30
    t_rte = P * F * A * L * 2
    bytes_in = (4 + (2 if cloudy else 0)) * 8 * P * L
    bytes_out = 6 * 8 * P * F * A + (8 * P * F * L if tau_layers else 0)
    return dict(w_abs=w_abs, w_pre=w_pre, t_pre=t_pre, w_rte=w_rte, t_rte=t_rte,
                flops_total=w_abs + w_pre + w_rte, bytes_in=bytes_in, bytes_out=bytes_out,
                # kernel-1 launch (absorption incl. the hoisted per-level work) and kernel-3 launch
                flops_k_abs=w_abs + w_pre, flops_k_rte=w_rte,
                # intermediates written by kernel 1 and read by kernel 3: awet, adry (P,F,L)
                bytes_k_abs=4 * 8 * P * L + (2 + (2 if cloudy else 0)) * 8 * P * F * L)


NOMINAL_FP64_FLOPS = 148 * 64 * 2 * 1.965e9  # 37.2 TFLOP/s (B200, non-tensor DFMA)
