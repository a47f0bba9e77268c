"""Drives the UNMODIFIED reference scripts staged in oracle/_ref/ (oracle/make_ref.py) for the timed reference arm.

TEST / BENCH INFRASTRUCTURE ONLY: imported by `bench.py --impl reference` (and by tests), never by the product.

The reference has no callable "one training step": its loop body lives inside train_lightgcn() (lightgcn_cu.py:601-656,
Version-2/lighgcn_cu_pop.py:820-869).  `ReferenceArm` therefore calls the reference's OWN functions and classes in the
order that loop does --

    build_cred_weighted_mats / build_message_passing_mats     graph build (NumPy + torch.sparse_coo .coalesce())
    edges_to_user_csr, sample_pos_item, sample_neg_item[_popmix]   the per-user Python sampler loop
    CredLightGCN.final_embeddings + score + l2_reg  |  LightGCN.get_user_item_emb + bpr_loss
    torch.optim.Adam(lr=cfg.lr): zero_grad, backward, step, loss.item()

-- and only the glue between those calls (the three `torch.tensor(list)` lines and, for lightgcn_cu.py, the four
lines that assemble the loss at CU:635-648) is restated here, with the line numbers it follows.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import pathlib
import time

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
MODULES = {"cu": "lightgcn_cu.py", "v2": "lighgcn_cu_pop.py", "da": "lightgcn_cu_pop_degree_aware.py"}


def available() -> bool:
    return all((REF_DIR / f).exists() for f in MODULES.values())


def load(variant: str, device: str = "cpu"):
    """Import one staged reference script as a module (its `main()` is guarded; import only builds `cfg`)."""
    path = REF_DIR / MODULES[variant]
    if not path.exists():
        raise FileNotFoundError(f"{path} is missing: run `python oracle/make_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(f"reference_{variant}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.cfg.device = device
    return mod


class ReferenceArm:
    """The reference's training step on (train_edges int32[2, E], cred float32[U]) with its own code."""

    def __init__(self, variant: str, train_edges: np.ndarray, num_users: int, num_items: int, cred: np.ndarray,
                 emb_dim: int, num_layers: int, device: str = "cpu", seed: int = 42, quiet: bool = True):
        self.variant, self.device = variant, device
        self.ref = ref = load(variant, device)
        ref.cfg.emb_dim, ref.cfg.num_layers, ref.cfg.seed = int(emb_dim), int(num_layers), int(seed)
        self.cfg = ref.cfg
        self.num_users, self.num_items = int(num_users), int(num_items)
        train_edges = np.ascontiguousarray(train_edges)
        sink = io.StringIO() if quiet else None
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            ref.set_seed(seed)                                         # CU:83-86 (np + torch)
            t0 = time.perf_counter()
            if variant == "cu":                                         # CU:580-586
                M_ui, M_iu, deg_i = ref.build_cred_weighted_mats(train_edges, num_users, num_items,
                                                                 np.asarray(cred, dtype=np.float32), device)
                pop = deg_i / max(float(deg_i.max()), 1.0)              # CU:583-584
                self.pop_t = torch.tensor(pop, dtype=torch.float32, device=device)
                self.model = ref.CredLightGCN(num_users, num_items, emb_dim, num_layers, M_ui, M_iu).to(device)
            else:                                                       # V2:787-792, DA:632-637
                cred_t = torch.tensor(np.asarray(cred, dtype=np.float32), device=device)
                M_ui, M_iu = ref.build_message_passing_mats(train_edges, num_users, num_items, cred_t, device=device)
                self.model = ref.LightGCN(num_users, num_items, emb_dim, num_layers, M_ui, M_iu).to(device)
            self.build_mats_s = time.perf_counter() - t0
            self.opt = torch.optim.Adam(self.model.parameters(), lr=ref.cfg.lr)     # CU:587, V2:793
            t0 = time.perf_counter()
            self.train_csr = ref.edges_to_user_csr(train_edges, num_users)          # CU:572
            self.build_csr_s = time.perf_counter() - t0
            self.pop_prob = None
            if variant == "v2":                                         # V2:805-810
                item_deg = np.bincount(train_edges[1].astype(np.int64), minlength=num_items).astype(np.float64)
                p = np.power(item_deg + 1.0, ref.cfg.neg_pop_gamma)
                self.pop_prob = (p / (p.sum() + 1e-12)).astype(np.float64)
        indptr = self.train_csr[0]
        self.train_users = np.where((indptr[1:] - indptr[:-1]) > 0)[0]              # CU:592
        self.rng = np.random.default_rng(ref.cfg.seed)                              # CU:589

    def sample_batch(self, batch_users):
        """The reference's per-user Python sampling loop for one batch (CU:611-629 / V2:835-856)."""
        ref, (indptr, indices) = self.ref, self.train_csr
        used, pos, neg = [], [], []
        for u in batch_users:
            p = ref.sample_pos_item(indptr, indices, int(u), self.rng)
            if p is None:
                continue
            if self.pop_prob is not None:
                n = ref.sample_neg_item_popmix(indptr, indices, int(u), self.num_items, self.rng,
                                               pop_prob=self.pop_prob, mix_pop=ref.cfg.neg_mix_pop,
                                               max_tries=ref.cfg.neg_max_tries)
            else:
                n = ref.sample_neg_item(indptr, indices, int(u), self.num_items, self.rng)
            used.append(int(u))
            pos.append(p)
            neg.append(n)
        dev = self.device
        return (torch.tensor(used, device=dev, dtype=torch.long), torch.tensor(pos, device=dev, dtype=torch.long),
                torch.tensor(neg, device=dev, dtype=torch.long))

    def step(self, users_t, pos_t, neg_t) -> float:
        """final embeddings + loss + zero_grad + backward + Adam step + loss.item() (CU:632-654 / V2:858-865)."""
        m, cfg = self.model, self.cfg
        if self.variant == "cu":
            e_u, e_i = m.final_embeddings()
            pos_scores = m.score(users_t, pos_t, e_u, e_i)
            neg_scores = m.score(users_t, neg_t, e_u, e_i)
            loss_bpr = -torch.log(torch.sigmoid(pos_scores - neg_scores) + 1e-12).mean()
            loss_fair = (self.pop_t[pos_t] * pos_scores).mean()
            loss_reg = m.l2_reg(users_t, pos_t, neg_t)
            loss = loss_bpr + cfg.lambda_fair * loss_fair + cfg.lambda_reg * loss_reg
        else:
            user_emb, item_emb = m.get_user_item_emb()
            loss = m.bpr_loss(users_t, pos_t, neg_t, user_emb, item_emb, cfg.reg)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.item())

    def batches(self, batch_size: int):
        """Shuffled user batches of one epoch (CU:603-608)."""
        users = self.train_users.copy()
        self.rng.shuffle(users)
        return [users[s:s + batch_size] for s in range(0, len(users), batch_size)]
