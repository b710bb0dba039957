"""Name -> class registry: the plugin boundary of the path (the reference's is sgmse/util/registry.py:5-34).

The reference surface is kept - ``register(name)`` as a class decorator that warns when a name is taken again,
``get_by_name`` raising ``ValueError`` for an unknown name, ``get_all_names`` - and extended with what this repo needs:
several spellings for one class (``register("ncsnpplarge", "ncsnpp-large")``), ``name in registry``, iteration in
registration order, and an error message that lists what IS registered.
"""
import warnings
from typing import Dict, Iterator, List, Tuple


class Registry:
    def __init__(self, managed_thing: str):
        self.managed_thing = managed_thing                 # used in warnings / errors, e.g. "Backbone"
        self._entries: Dict[str, type] = {}

    # ---- registration ------------------------------------------------------------------------------------------
    def _put(self, name: str, cls: type) -> None:
        if name in self._entries:
            warnings.warn(f"{self.managed_thing} with name '{name}' doubly registered, old class will be replaced.")
        self._entries[name] = cls

    def register(self, name: str, *aliases: str):
        """Class decorator: ``@registry.register("ncsnpp")`` (further positional names are aliases of the same class)."""
        names: Tuple[str, ...] = (name,) + aliases

        def bind(cls):
            for n in names:
                self._put(n, cls)
            return cls

        return bind

    # ---- lookup ------------------------------------------------------------------------------------------------
    def get_by_name(self, name: str):
        cls = self._entries.get(name)
        if cls is None:
            known = ", ".join(sorted(self._entries)) or "nothing registered"
            raise ValueError(f"{self.managed_thing} with name '{name}' unknown. (known: {known})")
        return cls

    def get_all_names(self) -> List[str]:
        return list(self._entries)

    def __contains__(self, name: str) -> bool:
        return name in self._entries

    def __iter__(self) -> Iterator[str]:
        return iter(self._entries)

    def __len__(self) -> int:
        return len(self._entries)
