"""Router.  Drop-in for RouterNetwork (expertsim/models/routers/router.py:6-26 of the reference): MLP 9-128-64-32-E
followed by a Gumbel-softmax; forward runs the fused gating kernel."""
import torch
import torch.nn.functional as F

from ... import _lib as L
from .._base import ArenaModule


class RouterNetwork(ArenaModule):
    ARCH, KIND = "proton", "router"

    def __init__(self, cond_dim, n_experts, **kwargs):
        super().__init__()
        self.name = "router-architecture-2"
        self.n_experts = n_experts
        if cond_dim != 9 or not 1 <= n_experts <= 16:
            raise ValueError("the fused gating kernel supports cond_dim=9 and 1..16 experts")
        self._init_params(dict(cond_dim=cond_dim, n_experts=n_experts), cond_dim=cond_dim, n_experts=n_experts)

    def raw_forward(self, cond, gumbel, tau):
        """-> dict(logits, gates, idx, h1, h2, h3, hist) on the device; gumbel is the injected -log(Exp(1)) noise."""
        a = self._home()
        B, E, dev = cond.shape[0], self.n_experts, cond.device
        nblk = (B + 255) // 256
        o = dict(logits=torch.empty(B, E, device=dev), gates=torch.empty(B, E, device=dev),
                 idx=torch.empty(B, dtype=torch.int64, device=dev), h1=torch.empty(B, 128, device=dev),
                 h2=torch.empty(B, 64, device=dev), h3=torch.empty(B, 32, device=dev),
                 hist=torch.empty(nblk, E, dtype=torch.int32, device=dev))
        L.call("es_router_fwd", cond, B, E, a.addr("fc_layers.0.weight"), a.addr("fc_layers.0.bias"), a.addr("fc_layers.2.weight"),
               a.addr("fc_layers.2.bias"), a.addr("fc_layers.4.weight"), a.addr("fc_layers.4.bias"), a.addr("fc_layers.6.weight"),
               a.addr("fc_layers.6.bias"), gumbel, float(tau), o["logits"], o["gates"], o["idx"], o["h1"], o["h2"], o["h3"], o["hist"])
        return o

    @torch.no_grad()
    def forward(self, cond, tau=1.0, hard=False):
        cond = cond.float().contiguous()
        gumbel = -torch.empty(cond.shape[0], self.n_experts, device=cond.device).exponential_().log()
        o = self.raw_forward(cond, gumbel, tau)
        gates = o["gates"]
        if hard:
            gates = F.one_hot(o["idx"], self.n_experts).to(gates.dtype)
        return gates, o["logits"]
