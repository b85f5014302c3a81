"""Transitive subclass enumeration (reference: mct_quantizers/common/get_all_subclasses.py:18-31)."""
from typing import Set


def get_all_subclasses(cls: type) -> Set[type]:
    found, stack = set(), list(cls.__subclasses__())
    while stack:
        sub = stack.pop()
        if sub not in found:
            found.add(sub)
            stack.extend(sub.__subclasses__())
    return found
