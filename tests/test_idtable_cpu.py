"""IdTable reproduces the reference's per-query dict inversion (hippocampal.py:312) incrementally."""
import random

from aura_snn_rag_b200 import IdTable


def test_owner_equals_dict_inversion_on_random_histories():
    rnd = random.Random(1)
    for _ in range(300):
        t, ref = IdTable(), {}
        for _ in range(80):
            mid, row = f"id{rnd.randint(0, 14)}", rnd.randint(0, 6)
            t.set(mid, row)
            ref[mid] = row
            inv = {v: k for k, v in ref.items()}
            assert t.id_to_idx == ref
            assert list(t.id_to_idx) == list(ref)          # same dict order
            for r in range(7):
                assert t.owner(r) == inv.get(r)


def test_full_bank_quirk_last_first_inserted_id_owns_row0():
    t = IdTable()
    for i in range(5):
        t.set(f"m{i}", i)
    for i in range(5, 9):          # bank full: every write lands on row 0 (hippocampal.py:200-202)
        t.set(f"m{i}", 0)
    assert t.owner(0) == "m8" and t.owner(1) == "m1"
    t.set("m6", 0)                 # rewriting an old key keeps its dict position -> m8 still owns row 0
    assert t.owner(0) == "m8"
