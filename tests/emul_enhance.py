"""Emulated enhance(): the product's host driver logic (tables, op lists) executed with the
PyTorch emulator on CPU.  Used by tests and for the precision-policy study quoted in DESIGN.md."""
import torch
import torch.nn.functional as F

from oracle import emulator as E
from open_universe_b200.engine import program as P


def emulated_enhance(model, oracle, mix, n_steps, noise, quant, epsilon=None):
    """mix (B,T); noise: list of unit-variance (B,1,T_pad).  Mirrors Universe.enhance."""
    epsilon = model.diff_kwargs.epsilon if epsilon is None else epsilon
    mix = mix[:, None, :]
    B, _, mix_len = mix.shape
    mixp, pad = model.pad(mix)
    mixn = oracle.normalize(mixp)
    t_pad = mixn.shape[-1]
    sigma, net_sigma, in_scale, coef, _ = model._sampler_tables(n_steps, epsilon, mixn)
    cp = P.lower_conditioner(model.condition_model, B, t_pad, need_signal_tail=False)
    bufs, _, _ = E.run_program(cp, {"x": mixn, "x_wav": mixn}, quant=quant)
    cond = [bufs[cp.outputs[f"cond{i}"]] for i in range(5)]
    net = model.get_score_model()
    sp = P.lower_score_network(net, B, t_pad)
    pp = P.lower_cond_projection(net, B, sp.meta["lengths"])
    pb, _, _ = E.run_program(pp, {f"cond{i}": c for i, c in enumerate(cond)}, quant=quant)
    from oracle.universe_oracle import sigma_embedding
    g = sigma_embedding(oracle.cfg["score_model"], oracle.sd, oracle.score_prefix + ".sigma_block",
                        torch.log10(net_sigma))
    film = E.film_table(sp, g)
    sc = {k: v for k, v in pb.items() if k.startswith("sc")}
    x = noise[0] * sigma[0]
    for n in range(n_steps):
        inputs = {"x": x}
        inputs.update(sc)
        z = noise[n + 1] * sigma[n + 1] if n < n_steps - 1 else None
        _, _, x = E.run_program(sp, inputs, film=film[n:n + 1].expand(B, -1),
                                in_scale=in_scale[n].expand(B), coef=coef[n][None].expand(B, 3),
                                noise=z, quant=quant)
    x = model.unpad(x, pad)
    scale = abs(x).max(dim=-1, keepdim=True).values
    x = torch.where(scale > 1.0, x / scale, x)
    return x[:, 0, :]
