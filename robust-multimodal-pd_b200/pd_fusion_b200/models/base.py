"""Model contract shared by every head (reference: models/base.py:4-20)."""
from abc import ABC, abstractmethod


class BaseModel(ABC):
    @abstractmethod
    def train(self, X, y, val_data=None):
        ...

    @abstractmethod
    def predict_proba(self, X, masks=None):
        ...

    @abstractmethod
    def save(self, path):
        ...

    @classmethod
    @abstractmethod
    def load(cls, path):
        ...
