"""Row <-> memory-id bookkeeping of the memory bank (host side, pure Python).

The reference keeps ``id_to_idx: Dict[str, int]`` (src/core/hippocampal.py:103,240) and, on EVERY
query, inverts the whole dict to map result rows back to ids (``{v: k for k, v in
id_to_idx.items()}``, :312) - O(M) Python per query, 103 ms at 1M memories.
This table maintains that inverse incrementally and reproduces the dict inversion exactly,
including its corner cases:

* several ids can point at one row (full bank: every write lands on row 0, :200-202); the
  inversion keeps the id that comes LAST in dict order, and dict order is the order in which
  keys were FIRST inserted (re-assigning an existing key keeps its position);
* an id written again to another row leaves its old row without an owner unless another id
  still points there; rows without an owner are dropped from results (:316).
"""
from __future__ import annotations

from typing import Dict, Optional, Union


class IdTable:
    def __init__(self) -> None:
        self.id_to_idx: Dict[str, int] = {}          # the reference's public attribute
        self._pos: Dict[str, int] = {}               # first-insertion position of every id
        self._ids_of_row: Dict[int, Union[str, set]] = {}   # one id (str) or several (set)
        self._owner: Dict[int, str] = {}             # row -> id the dict inversion would give

    def __len__(self) -> int:
        return len(self.id_to_idx)

    def set(self, memory_id: str, row: int) -> None:
        """id_to_idx[memory_id] = row (hippocampal.py:240)."""
        old = self.id_to_idx.get(memory_id)
        if memory_id not in self._pos:
            self._pos[memory_id] = len(self._pos)
        self.id_to_idx[memory_id] = row
        if old is not None and old != row:
            self._detach(memory_id, old)
        if old == row:
            return
        cur = self._ids_of_row.get(row)
        if cur is None:
            self._ids_of_row[row] = memory_id
            self._owner[row] = memory_id
            return
        if isinstance(cur, str):
            cur = {cur}
            self._ids_of_row[row] = cur
        cur.add(memory_id)
        if self._pos[memory_id] > self._pos[self._owner[row]]:
            self._owner[row] = memory_id

    def _detach(self, memory_id: str, row: int) -> None:
        cur = self._ids_of_row.get(row)
        if isinstance(cur, set):
            cur.discard(memory_id)
            if not cur:
                del self._ids_of_row[row]
                del self._owner[row]
            elif self._owner[row] == memory_id:
                self._owner[row] = max(cur, key=self._pos.__getitem__)
        elif cur == memory_id:
            del self._ids_of_row[row]
            del self._owner[row]

    def owner(self, row: int) -> Optional[str]:
        """The id ``{v: k for k, v in id_to_idx.items()}[row]`` would hold, or None."""
        return self._owner.get(row)
