// cache.cu -- HBM residency cache: what VICacheManager is for CPU indexes (reference:
// VectorIndex/Cache/VICacheManager.cpp:65-157 over ClickHouse's LRUResourceCache, keyed by CacheKey,
// VectorIndex/Cache/VICacheObject.h:119-137, weighted by getResourceUsage().memory_usage_bytes).  Host-only logic;
// the objects it owns are the device-resident handles of this library.  SURVEY section 8(f)3.
#include <list>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"

using namespace b200;

namespace {

struct Entry {
    std::string key;
    int kind = B200_CACHE_OPAQUE;
    void *handle = nullptr;
    uint64_t bytes = 0;
    int pins = 0;
    bool expired = false;  // removed from the map while pinned: freed at the last release
    void (*deleter)(void *) = nullptr;
};

struct Cache {
    std::mutex mu;
    uint64_t capacity = ~0ull;  // unlimited until VICacheManager::setCacheSize is mirrored
    uint64_t used = 0, hits = 0, misses = 0, evictions = 0;
    std::list<Entry> lru;                                            // front = most recently used
    std::unordered_map<std::string, std::list<Entry>::iterator> map;  // live (not expired) entries
    std::list<Entry> zombies;                                        // expired but still pinned
};

Cache &cache() {
    static Cache c;
    return c;
}

void destroy(Entry &e) {
    if (!e.handle) return;
    switch (e.kind) {
        case B200_CACHE_CORPUS: b200_corpus_free(reinterpret_cast<b200_corpus *>(e.handle)); break;
        case B200_CACHE_INDEX: b200_index_free(reinterpret_cast<b200_index *>(e.handle)); break;
        case B200_CACHE_BM25: b200_bm25_free(reinterpret_cast<b200_bm25 *>(e.handle)); break;
        default:
            if (e.deleter) e.deleter(e.handle);
    }
    e.handle = nullptr;
}

// evict unpinned entries, least recently used first, until `need` more bytes fit; false if they cannot
bool make_room(Cache &c, uint64_t need, std::list<Entry> &doomed) {
    if (need > c.capacity) return false;
    auto it = c.lru.end();
    while (c.used + need > c.capacity && it != c.lru.begin()) {
        --it;
        if (it->pins > 0) continue;
        c.used -= it->bytes;
        c.evictions++;
        c.map.erase(it->key);
        auto victim = it++;
        doomed.splice(doomed.end(), c.lru, victim);
    }
    return c.used + need <= c.capacity;
}

int put_impl(const char *key, int kind, void *handle, uint64_t bytes, void (*deleter)(void *), void **resident) {
    if (!key || !handle || !resident || kind < B200_CACHE_CORPUS || kind > B200_CACHE_OPAQUE)
        return fail(B200_ERR_INVALID, "b200_cache_put: bad arguments");
    std::list<Entry> doomed;  // destroyed outside the lock: freeing device memory synchronises
    int rc = B200_OK;
    {
        Cache &c = cache();
        std::lock_guard<std::mutex> lk(c.mu);
        auto f = c.map.find(key);
        if (f != c.map.end()) {  // getOrSet: the resident object wins
            f->second->pins++;
            c.lru.splice(c.lru.begin(), c.lru, f->second);
            *resident = f->second->handle;
            c.hits++;
        } else if (!make_room(c, bytes, doomed)) {
            rc = fail(B200_ERR_NOMEM, "cache: object does not fit next to the pinned entries");
        } else {
            Entry e;
            e.key = key;
            e.kind = kind;
            e.handle = handle;
            e.bytes = bytes;
            e.pins = 1;
            e.deleter = deleter;
            c.lru.push_front(e);
            c.map[e.key] = c.lru.begin();
            c.used += bytes;
            *resident = handle;
        }
    }
    for (auto &e : doomed) destroy(e);
    return rc;
}

}  // namespace

extern "C" int b200_cache_set_capacity(uint64_t bytes) {
    std::list<Entry> doomed;
    {
        Cache &c = cache();
        std::lock_guard<std::mutex> lk(c.mu);
        c.capacity = bytes;
        make_room(c, 0, doomed);
    }
    for (auto &e : doomed) destroy(e);
    return B200_OK;
}

extern "C" int b200_cache_get(const char *key, void **handle, int *kind) {
    if (!key || !handle) return fail(B200_ERR_INVALID, "b200_cache_get: bad arguments");
    Cache &c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto f = c.map.find(key);
    if (f == c.map.end()) {
        c.misses++;
        *handle = nullptr;
        return fail(B200_ERR_NOT_FOUND, std::string("cache miss: ") + key);
    }
    c.hits++;
    f->second->pins++;
    c.lru.splice(c.lru.begin(), c.lru, f->second);
    *handle = f->second->handle;
    if (kind) *kind = f->second->kind;
    return B200_OK;
}

extern "C" int b200_cache_put(const char *key, int kind, void *handle, uint64_t bytes, void **resident) {
    if (kind == B200_CACHE_OPAQUE) return fail(B200_ERR_INVALID, "use b200_cache_put_opaque for caller-defined objects");
    return put_impl(key, kind, handle, bytes, nullptr, resident);
}

extern "C" int b200_cache_put_opaque(const char *key, void *handle, uint64_t bytes, void (*deleter)(void *), void **resident) {
    return put_impl(key, B200_CACHE_OPAQUE, handle, bytes, deleter, resident);
}

// One pin is dropped from the entry that holds `handle` under `key`: the live entry if it is that generation, else the
// expired-but-still-pinned one.  handle == nullptr keeps the round-1 behaviour (live entry first) for callers that never
// re-put a key while an older generation is still held.  The reference's LRUResourceCache hands out shared_ptr holders,
// so a release can never hit another generation there; matching on the handle gives the same guarantee here.
static int release_impl(const char *key, const void *handle) {
    if (!key) return fail(B200_ERR_INVALID, "b200_cache_release: bad arguments");
    std::list<Entry> doomed;
    int rc = B200_OK;
    {
        Cache &c = cache();
        std::lock_guard<std::mutex> lk(c.mu);
        auto f = c.map.find(key);
        if (f != c.map.end() && f->second->pins > 0 && (!handle || f->second->handle == handle)) {
            f->second->pins--;
        } else {
            // an expired entry keeps its pins in the zombie list (oldest first)
            auto z = c.zombies.begin();
            while (z != c.zombies.end() && !(z->key == key && (!handle || z->handle == handle))) ++z;
            if (z == c.zombies.end()) {
                rc = fail(B200_ERR_INVALID, std::string("cache release without a pin: ") + key);
            } else if (--z->pins == 0) {
                doomed.splice(doomed.end(), c.zombies, z);
            }
        }
        if (rc == B200_OK) make_room(c, 0, doomed);  // a shrunk capacity may now be reachable
    }
    for (auto &e : doomed) destroy(e);
    return rc;
}

extern "C" int b200_cache_release(const char *key) { return release_impl(key, nullptr); }

extern "C" int b200_cache_release_handle(const char *key, const void *handle) {
    if (!handle) return fail(B200_ERR_INVALID, "b200_cache_release_handle: null handle");
    return release_impl(key, handle);
}

static void expire_locked(Cache &c, std::list<Entry>::iterator it, std::list<Entry> &doomed) {
    c.used -= it->bytes;
    c.map.erase(it->key);
    if (it->pins > 0) {
        it->expired = true;
        c.zombies.splice(c.zombies.end(), c.lru, it);
    } else {
        doomed.splice(doomed.end(), c.lru, it);
    }
}

extern "C" int b200_cache_expire(const char *key) {
    if (!key) return fail(B200_ERR_INVALID, "b200_cache_expire: bad arguments");
    std::list<Entry> doomed;
    int rc = B200_OK;
    {
        Cache &c = cache();
        std::lock_guard<std::mutex> lk(c.mu);
        auto f = c.map.find(key);
        if (f == c.map.end()) rc = fail(B200_ERR_NOT_FOUND, std::string("cache expire: no such key: ") + key);
        else expire_locked(c, f->second, doomed);
    }
    for (auto &e : doomed) destroy(e);
    return rc;
}

extern "C" int b200_cache_expire_prefix(const char *prefix, int64_t *out_removed) {
    if (!prefix) return fail(B200_ERR_INVALID, "b200_cache_expire_prefix: bad arguments");
    const std::string pre(prefix);
    std::list<Entry> doomed;
    int64_t removed = 0;
    {
        Cache &c = cache();
        std::lock_guard<std::mutex> lk(c.mu);
        for (auto it = c.lru.begin(); it != c.lru.end();) {
            auto cur = it++;
            if (cur->key.compare(0, pre.size(), pre) == 0) {
                expire_locked(c, cur, doomed);
                removed++;
            }
        }
    }
    for (auto &e : doomed) destroy(e);
    if (out_removed) *out_removed = removed;
    return B200_OK;
}

extern "C" int b200_cache_stats(uint64_t *capacity, uint64_t *used, uint64_t *items, uint64_t *hits, uint64_t *misses,
                                uint64_t *evictions) {
    Cache &c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    if (capacity) *capacity = c.capacity;
    if (used) *used = c.used;
    if (items) *items = c.map.size();
    if (hits) *hits = c.hits;
    if (misses) *misses = c.misses;
    if (evictions) *evictions = c.evictions;
    return B200_OK;
}
