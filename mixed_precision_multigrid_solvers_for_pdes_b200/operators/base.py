"""Operator protocol of the reference (operators/base.py:11-55): ``apply`` + ``can_apply``."""
from abc import ABC, abstractmethod


class BaseOperator(ABC):
    def __init__(self, name: str = "BaseOperator"):
        self.name = name

    @abstractmethod
    def apply(self, grid, field):
        ...

    @abstractmethod
    def can_apply(self, grid) -> bool:
        ...

    def __str__(self) -> str:
        return self.name

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(name='{self.name}')"
