"""HBM residency cache (device-side VICacheManager, include/b200_search.h): LRU-by-bytes semantics with pinning.
The bookkeeping is host code, so it is tested on CPU with caller-defined (opaque) objects; tests/test_gpu_edges.py
puts real device corpora through it."""
import ctypes as C

import pytest

from myscaledb_b200 import search as S


@pytest.fixture()
def tracker():
    freed = []
    cb = S._DELETER(lambda h: freed.append(int(h)))
    S.cache_expire_prefix("t/")
    S.cache_set_capacity(1000)
    yield freed, cb
    S.cache_expire_prefix("t/")
    S.cache_set_capacity(2 ** 63)


def test_lru_by_bytes_evicts_least_recently_used_unpinned(tracker):
    freed, cb = tracker
    for i, w in enumerate((400, 400)):
        assert S.cache_put_opaque(f"t/p{i}", 100 + i, w, cb) == 100 + i
        S.cache_release(f"t/p{i}")
    h, kind = S.cache_get("t/p0")          # touch p0: p1 becomes the LRU entry
    assert (h, kind) == (100, S.CACHE_OPAQUE)
    S.cache_release("t/p0")
    S.cache_put_opaque("t/p2", 102, 400, cb)  # 1200 > 1000: evicts p1, not p0
    S.cache_release("t/p2")
    assert freed == [101]
    with pytest.raises(S.CacheMiss):
        S.cache_get("t/p1")
    st = S.cache_stats()
    assert st["used"] == 800 and st["items"] >= 2 and st["evictions"] >= 1


def test_pinned_entries_are_never_evicted_and_put_fails_when_nothing_fits(tracker):
    freed, cb = tracker
    S.cache_put_opaque("t/a", 1, 600, cb)       # stays pinned
    with pytest.raises(S.B200Error) as ei:
        S.cache_put_opaque("t/b", 2, 600, cb)   # would need to evict the pinned entry
    assert ei.value.code == 5 and freed == []   # B200_ERR_NOMEM, ownership of 2 stays with the caller
    S.cache_release("t/a")
    S.cache_put_opaque("t/b", 2, 600, cb)       # now a is evictable
    assert freed == [1]
    S.cache_release("t/b")
    with pytest.raises(S.B200Error):
        S.cache_put_opaque("t/huge", 3, 5000, cb)  # larger than the whole cache


def test_get_or_set_keeps_the_resident_object(tracker):
    freed, cb = tracker
    assert S.cache_put_opaque("t/k", 7, 100, cb) == 7
    assert S.cache_put_opaque("t/k", 8, 100, cb) == 7   # already resident: the second object stays the caller's
    S.cache_release("t/k"); S.cache_release("t/k")
    assert freed == []
    with pytest.raises(S.B200Error):
        S.cache_release("t/k")                          # no pin left


def test_force_expire_defers_the_free_until_the_last_release(tracker):
    freed, cb = tracker
    S.cache_put_opaque("t/x", 11, 100, cb)
    S.cache_expire("t/x")                # VICacheManager::forceExpire while a search still holds the index
    assert freed == []
    with pytest.raises(S.CacheMiss):
        S.cache_get("t/x")               # gone for new readers
    S.cache_put_opaque("t/x", 12, 100, cb)   # a rebuilt index can be cached under the same key meanwhile
    S.cache_release("t/x")               # releases the live entry's pin first
    S.cache_release("t/x")               # then the expired one: freed now
    assert freed == [11]
    with pytest.raises(S.CacheMiss):
        S.cache_expire("t/nope")


def test_shrinking_capacity_and_prefix_expiry(tracker):
    freed, cb = tracker
    for i in range(5):
        S.cache_put_opaque(f"t/db/tbl/part{i}", 20 + i, 150, cb)
        S.cache_release(f"t/db/tbl/part{i}")
    S.cache_set_capacity(400)            # updateMaxWeight: evicts LRU-first down to 2 entries
    assert freed == [20, 21, 22]
    assert S.cache_expire_prefix("t/db/tbl/") == 2   # dropped table
    assert sorted(freed) == [20, 21, 22, 23, 24]
    assert S.cache_stats()["used"] == 0


@pytest.mark.parametrize("old_first", [True, False])
def test_release_by_handle_never_unpins_another_generation(tracker, old_first):
    """key K is expired while pinned (zombie), then K is put again (index rebuilt under the same CacheKey).
    The old holder's release must drop the OLD generation's pin, in either order -- otherwise the new entry reaches 0 pins
    while its holder still uses it and make_room frees it under the holder's feet."""
    freed, cb = tracker
    assert S.cache_put_opaque("t/K", 11, 600, cb) == 11        # generation 1, pinned by holder A
    S.cache_expire("t/K")                                       # expired while pinned -> zombie
    assert freed == []
    assert S.cache_put_opaque("t/K", 12, 600, cb) == 12        # generation 2, pinned by holder B
    if old_first:
        S.cache_release("t/K", handle=11)
        assert freed == [11]                                    # the zombie goes at its last release
        # B's entry is still pinned: something that needs the room must NOT evict it
        with pytest.raises(S.B200Error):
            S.cache_put_opaque("t/other", 13, 600, cb)
        assert freed == [11]
        S.cache_release("t/K", handle=12)
    else:
        S.cache_release("t/K", handle=12)                       # B is done first; A still holds generation 1
        assert freed == []
        S.cache_put_opaque("t/other", 13, 600, cb)              # evicts the unpinned generation 2
        assert freed == [12]
        S.cache_release("t/other")
        S.cache_release("t/K", handle=11)
        assert freed == [12, 11]
    with pytest.raises(S.B200Error):
        S.cache_release("t/K", handle=11)                       # no pin of that generation left
