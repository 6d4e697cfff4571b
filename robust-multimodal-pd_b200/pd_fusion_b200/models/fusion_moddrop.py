"""Fusion-ModDrop (reference: models/fusion_moddrop.py).

`ModalityDropoutNet` keeps the reference's layout (feature blocks in sorted(modality) order, `net.{0,3,6,..}`
keys).  `predict_proba` runs on the device through pdf_moddrop_sweep; `predict_proba_sweep` evaluates every
scenario mask of a missingness sweep in ONE launch (layer-1 partials per modality computed once).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from ..heads import ModDropSweep
from ..utils.torch_utils import get_torch_device
from .base import BaseModel


class ModalityDropoutNet(nn.Module):
    def __init__(self, modality_dims: Dict[str, int], hidden_dims: List[int], dropout: float = 0.2):
        super().__init__()
        self.modality_dims = modality_dims
        self.mod_names = sorted(modality_dims)
        self.slices, pos = {}, 0
        for mod in self.mod_names:
            self.slices[mod] = (pos, pos + modality_dims[mod])
            pos += modality_dims[mod]
        layers, width = [], pos
        for h in hidden_dims:
            layers += [nn.Linear(width, h), nn.ReLU(), nn.Dropout(dropout)]
            width = h
        layers += [nn.Linear(width, 1), nn.Sigmoid()]
        self.net = nn.Sequential(*layers)

    def forward(self, x, training_dropout: bool = False, drop_rate: float = 0.0):
        if training_dropout and self.training:
            keep = torch.ones_like(x)
            for mod in self.mod_names:          # one np.random.rand() per modality per batch, as the reference draws them
                if np.random.rand() < drop_rate:
                    a, b = self.slices[mod]
                    keep[:, a:b] = 0.0
            x = x * keep
        return self.net(x)


class ModalityDropoutModel(BaseModel):
    def __init__(self, modality_dims, params):
        self.params = params
        self.modality_dims = modality_dims
        self.model = ModalityDropoutNet(modality_dims, params["hidden_dims"], params.get("dropout", 0.2))
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=params["lr"], weight_decay=params.get("weight_decay", 0.0))
        self.criterion = nn.BCELoss()
        self._sweep: Optional[ModDropSweep] = None

    def train(self, X, y, val_data=None):
        """Host-side torch training (SURVEY.md 8f rank 2 'next'); modality dropout draws follow the reference."""
        Xt = torch.as_tensor(np.asarray(X), dtype=torch.float32)
        yt = torch.as_tensor(np.asarray(y), dtype=torch.float32).view(-1, 1)
        rate, bs = self.params.get("moddrop_rate", 0.2), self.params.get("batch_size", 32)
        for _ in range(self.params["epochs"]):
            self.model.train()
            order = torch.randperm(len(Xt))
            for i in range(0, len(Xt), bs):
                sel = order[i:i + bs]
                self.optimizer.zero_grad()
                loss = self.criterion(self.model(Xt[sel], training_dropout=True, drop_rate=rate), yt[sel])
                loss.backward()
                self.optimizer.step()
        self._sweep = None

    def invalidate(self):
        self._sweep = None

    def _get_sweep(self) -> ModDropSweep:
        if self._sweep is None:
            self._sweep = ModDropSweep(self.model.state_dict(), self.modality_dims, device=get_torch_device())
        return self._sweep

    def predict_proba_sweep(self, X: np.ndarray, masks_snm: np.ndarray, order: List[str]) -> np.ndarray:
        """X [N,F] (unmasked); masks_snm uint8 [S,N,len(order)] -> probabilities [S,N] f32, one launch."""
        sw = self._get_sweep()
        cols = [order.index(m) if m in order else -1 for m in sw.mods]
        mk = np.ones(masks_snm.shape[:2] + (len(sw.mods),), dtype=np.uint8)
        for j, c in enumerate(cols):
            if c >= 0:
                mk[:, :, j] = masks_snm[:, :, c]
        dev = sw.device
        prob = sw.forward(torch.as_tensor(np.asarray(X), dtype=torch.float32).to(dev), torch.from_numpy(mk).to(dev))
        return prob.cpu().numpy()

    def predict_proba(self, X, masks=None):
        X = np.asarray(X)
        sw = self._get_sweep()
        mk = np.ones((1, X.shape[0], len(sw.mods)), dtype=np.uint8)
        if masks is not None:
            for j, mod in enumerate(sw.mods):
                if mod in masks:
                    # the reference multiplies by the float mask value; masks are {0,1} on every call site
                    mk[0, :, j] = np.asarray(masks[mod]) != 0
        return self.predict_proba_sweep(X, mk, sw.mods)[0]

    def save(self, path):
        torch.save(self.model.state_dict(), path)

    @classmethod
    def load(cls, path, modality_dims, params):
        """The reference's `load` is an empty stub (fusion_moddrop.py:119-123); this one restores the state_dict."""
        inst = cls(modality_dims, params)
        inst.model.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))
        return inst
