"""Name -> callable registry with the same surface as the reference's ``FunctionRegister``
(apps/api/src/register/__init__.py:8-143): decorator registration (bare or keyed, ``available`` flag,
``overwrite``), ``get`` / ``call`` / ``set_default`` / ``all_available`` and the same error behaviour
(duplicate key -> KeyError, unknown key -> KeyError, unavailable -> RuntimeError("Function '<k>' is not available.")).

Inside the reference server the "b200" backend registers into the reference's own registry object
(see INTEGRATION.md); this mirror exists so the backend, its tests and bench.py run without the reference.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterator, Optional


class FunctionRegister:
    def __init__(self, *, allow_overwrite: bool = False) -> None:
        self._entries: Dict[str, Dict[str, Any]] = {}
        self._allow_overwrite = allow_overwrite

    # -- registration ---------------------------------------------------------------------------
    def __call__(self, key_or_func=None, *, overwrite: Optional[bool] = None, available: bool = True):
        if callable(key_or_func) and overwrite is None:  # bare @registry
            self._add(key_or_func.__name__, key_or_func, self._allow_overwrite, available)
            return key_or_func
        key = key_or_func

        def decorate(func: Callable) -> Callable:
            self._add(key, func, self._allow_overwrite if overwrite is None else overwrite, available)
            return func

        return decorate

    def _add(self, key: str, func: Callable, allow: bool, available: bool) -> None:
        if key in self._entries and not allow:
            raise KeyError(f"Key '{key}' already registered. Use overwrite=True to replace.")
        self._entries[key] = {"func": func, "available": available}

    # -- lookup ---------------------------------------------------------------------------------
    def get(self, key: str) -> Callable:
        if key not in self._entries:
            raise KeyError(f"Key '{key}' not found in registry.")
        return self._entries[key]["func"]

    def is_available(self, key: str) -> bool:
        return key in self._entries and bool(self._entries[key]["available"])

    def set_availability(self, key: str, available: bool) -> None:
        if key not in self._entries:
            raise KeyError(f"Key '{key}' not found in registry.")
        self._entries[key]["available"] = available

    def call(self, *args, key: Optional[str] = None, **kwargs):
        if key is None and hasattr(self, "_default"):
            key = self._default
        if not self.is_available(key):
            raise RuntimeError(f"Function '{key}' is not available.")
        return self.get(key)(*args, **kwargs)

    def all(self) -> Dict[str, Callable]:
        return {k: e["func"] for k, e in self._entries.items()}

    def all_available(self) -> Dict[str, Callable]:
        return {k: e["func"] for k, e in self._entries.items() if e["available"]}

    def set_default(self, key: str) -> None:
        self._default = key

    def get_default(self) -> str:
        return self._default

    __getitem__ = get

    def __iter__(self) -> Iterator[str]:
        return iter(self._entries)

    def __len__(self) -> int:
        return len(self._entries)
