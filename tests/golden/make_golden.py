#!/usr/bin/env python3
"""Extract the reference's golden outputs for the ANN / BM25 hot path into JSON fixtures.

Run ONCE in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It copies *data only* (expected stdout rows of the reference's own SQL tests and
the literal inputs in their SQL) from
/root/reference/tests/queries/2_vector_search/*.reference|.sql|.sh into
tests/golden/*.json.  The tests never read /root/reference at run time
(it does not exist on the GPU box).
"""
import json
import os
import re

REF = "/root/reference/tests/queries/2_vector_search"
OUT = os.path.dirname(os.path.abspath(__file__))


def read(name):
    with open(os.path.join(REF, name)) as f:
        return f.read()


def rows(text):
    return [ln.split("\t") for ln in text.strip("\n").split("\n")]


def sections(text, is_header):
    """Split a .reference into {header: [rows]} using a header predicate."""
    out, cur = {}, None
    for ln in text.strip("\n").split("\n"):
        if is_header(ln):
            cur = ln
            out[cur] = []
        elif cur is not None:
            out[cur].append(ln.split("\t"))
    return out


def tuple_col(s):
    a, b = s.strip("()").split(",")
    return int(a), float(b)


def main():
    g = {}

    # 00001: FLAT index, 100 rows [n,n,n], query [0.1]*3, L2, top 10 (twice: before/after DETACH/ATTACH)
    r = rows(read("00001_mqvs_distance.reference"))
    g["00001_flat_l2"] = {
        "source": "tests/queries/2_vector_search/00001_mqvs_distance.{sh,reference}",
        "corpus": "row n = [n,n,n] for n in range(100)", "query": [0.1, 0.1, 0.1], "metric": "L2", "k": 10,
        "expect": [[int(x[0]), float(x[2])] for x in r[:10]],
        "expect_after_reload": [[int(x[0]), float(x[2])] for x in r[10:20]],
    }

    # 00012: brute force, 10030 rows? -> helper 00000_prepare_index_2.sh defines the corpus
    helper2 = read("helpers/00000_prepare_index_2.sh")
    r = rows(read("00012_mqvs_brute_force_search.reference"))
    g["00012_bruteforce_l2"] = {
        "source": "tests/queries/2_vector_search/00012_mqvs_brute_force_search.{sh,reference}",
        "helper_sql": [ln for ln in helper2.split("\n") if "INSERT" in ln or "CREATE" in ln],
        "query": [10020.1] * 3, "metric": "L2", "k": 100,
        "expect": [[int(x[0]), float(x[2])] for x in r],
    }

    # 00002: batch_distance L2 and IP, 2 parts of 50 rows, 3 queries, LIMIT 10 BY query
    s = sections(read("00002_mqvs_batch_distance.reference"), lambda ln: ln.startswith("-- "))
    for key, name in (("-- batch_distance of metric_type=L2", "00002_batch_l2"), ("-- batch_distance of metric_type=IP", "00002_batch_ip")):
        g[name] = {
            "source": "tests/queries/2_vector_search/00002_mqvs_batch_distance.{sh,reference}",
            "corpus": "two parts: rows [n,n,n] n in 0..49 and 50..99", "queries": [[0.1] * 3, [0.2] * 3, [50.1] * 3],
            "metric": name.split("_")[-1].upper(), "k": 10,
            "expect": [[int(x[0]), *tuple_col(x[2])] for x in s[key]],
        }

    # 00014: cosine brute force
    r = rows(read("00014_mqvs_distance_cosine_bruteforce.reference"))
    g["00014_cosine_bruteforce"] = {
        "source": "tests/queries/2_vector_search/00014_mqvs_distance_cosine_bruteforce.{sql,reference}",
        "corpus": "row n = [n, n+3, n+1] for n in range(1000)", "query": [8.0, 11.0, 9.0], "metric": "COSINE", "k": 5,
        "expect": [[int(x[0]), float(x[1])] for x in r],
    }

    # 00038: binary vectors (Hamming / Jaccard), brute force, batch, filter
    s = sections(read("00038_mqvs_binary_vector_feature.reference"), lambda ln: ln.startswith("-- "))
    b = {"source": "tests/queries/2_vector_search/00038_mqvs_binary_vector_feature.{sql,reference}",
         "corpus": "row n = bytes([n%256]*4) for n in range(1024)",
         "query": [100, 101, 102, 103],
         "batch_queries": [[0x55] * 4, [0, 255, 1, 254], [255] * 4],
         "filter": "100 < id < 120"}
    for hdr, key in (("-- Brute Force (Hamming)", "hamming_brute"), ("-- Search with filter (Hamming)", "hamming_filter"),
                     ("-- Brute Force (Jaccard)", "jaccard_brute"), ("-- Search with filter (Jaccard)", "jaccard_filter")):
        b[key] = [[int(x[0]), float(x[1])] for x in s[hdr]]
    for hdr, key in (("-- Batch distance (Hamming)", "hamming_batch"), ("-- Batch distance (Jaccard)", "jaccard_batch")):
        b[key] = [[int(x[0]), *tuple_col(x[1])] for x in s[hdr]]
    # LWD section: after DELETE WHERE id < 200, Hamming top-10
    lwd = [x for x in s["-- LWD"] if len(x) == 2 and x[0].isdigit()]
    b["hamming_after_lwd_lt200"] = [[int(x[0]), float(x[1])] for x in lwd[:10]]
    g["00038_binary"] = b

    # 00028: MSTG on 1000 x 768; query literal from SQL
    sql = read("00028_mqvs_index_mstg_build_search.sql")
    m = re.search(r"distance\(vector, \[([0-9eE+\-., ]+)\]\)", sql)
    q = [float(t) for t in m.group(1).split(",")]
    assert len(q) == 768, len(q)
    ref = read("00028_mqvs_index_mstg_build_search.reference").strip("\n").split("\n")
    num = [ln.split("\t") for ln in ref if re.match(r"^\d+\t[0-9.]+$", ln)]
    g["00028_mstg_768"] = {
        "source": "tests/queries/2_vector_search/00028_mqvs_index_mstg_build_search.{sql,reference}",
        "corpus": "row n, dim x: 0.00001*(n*768+x+1)*(-1 if x%2==0 else 1), n in range(1000), x in range(768)",
        "query": q, "k": 5,
        "expect_l2": [[int(x[0]), float(x[1])] for x in num[:5]],
        "expect_cosine": [[int(x[0]), float(x[1])] for x in num[5:10]],
        "expect_cosine_where_id_gt0": [[int(x[0]), float(x[1])] for x in num[10:15]],
        "expect_cosine_after_delete_id2": [[int(x[0]), float(x[1])] for x in num[15:20]],
        "sql_tail": [ln for ln in sql.split("\n") if ln.startswith(("SELECT id", "DELETE", "ALTER"))][:12]
        if False else None,
    }

    # 00035: ties (ids 0 and 2 both at 16)
    s = sections(read("00035_mqvs_two_stage_search.reference"), lambda ln: not re.match(r"^\d+(\t|$)", ln))
    hdr = "two stage search with MSTG type and min_bytes_to_build_vector_index=0"
    g["00035_ties"] = {
        "source": "tests/queries/2_vector_search/00035_mqvs_two_stage_search.{sql,reference}",
        "corpus": "row n = [n]*16 for n in range(1001) if n != 1", "query": [1.0] * 16, "metric": "L2", "k": 10,
        "filter": "id < 11",
        "expect_filtered": [[int(x[0]), float(x[1])] for x in s[hdr] if len(x) == 2],
        "expect_unfiltered": [[int(x[0]), float(x[1])] for x in s["disable two stage search"] if len(x) == 2],
    }

    # 00040 / 00041: BM25 + fusion
    sql = read("00040_mqvs_hybrid_search.sql")
    docs20 = re.findall(r"\((\d+), ?\[(\d+),\d+,\d+\], ?'((?:[^']|'')*)'\)", sql.split("SELECT 'support only one")[0])
    docs20 = [[int(a), int(b), c.replace("''", "'")] for a, b, c in docs20]
    assert len(docs20) == 20, len(docs20)
    arr_sql = sql.split("INSERT INTO t_vector_invert_array VALUES")[1].split(";\n")[0]
    arr_docs = []
    for mm in re.finditer(r"\((\d+), ?\[\d+,\d+,\d+\], ?\[((?:'(?:[^']|'')*',? ?)+)\]\)", arr_sql):
        pieces = [p.replace("''", "'") for p in re.findall(r"'((?:[^']|'')*)'", mm.group(2))]
        arr_docs.append([int(mm.group(1)), pieces])
    assert len(arr_docs) == 10, len(arr_docs)
    map_sql = sql.split("INSERT INTO t_vector_invert_map VALUES")[1].split(";\n")[0]
    map_docs = [[int(a), b] for a, b in re.findall(r"\((\d+), ?\[\d+,\d+,\d+\], ?\{'([^']*)':", map_sql)]
    assert len(map_docs) == 20, len(map_docs)
    multi_sql = sql.split("INSERT INTO t_vector_invert_multi VALUES")[1].split(";\n")[0]
    multi_docs = [[int(a), b.replace("''", "'"), c.replace("''", "'")] for a, b, c in
                  re.findall(r"\((\d+), ?\[\d+,\d+,\d+\], ?'((?:[^']|'')*)', ?'((?:[^']|'')*)'\)", multi_sql)]
    assert len(multi_docs) == 20, len(multi_docs)
    s = sections(read("00040_mqvs_hybrid_search.reference"), lambda ln: not re.match(r"^\d+(\t|$)", ln))

    def fl(h):
        return [[int(x[0]), float(x[1])] for x in s[h] if len(x) == 2]
    g["00040_hybrid"] = {
        "source": "tests/queries/2_vector_search/00040_mqvs_hybrid_search.{sql,reference}",
        "docs": docs20, "vector": "row id -> [id,id,id]", "query_vector": [1.0, 1.0, 1.0], "query_text": "Ancient",
        "text_search": fl("text search"),
        "text_search_where_id_lt10": fl("text search with WHERE clause"),
        "rsf": fl("hybrid search with relative score fusion"),
        "rrf": fl("hybrid search with rank fusion"),
        "rsf_where_id_lt10": fl("hybrid search rsf with WHERE clause")[:5],
        "array_docs": arr_docs, "array_query": "Military Strategy", "array_text_search": fl("text search on Array"),
        "map_docs": map_docs, "map_query": "Comics", "map_text_search": fl("text search on Map"),
        "multi_docs": multi_docs, "multi_doc2_query": "cultural", "multi_rsf_doc2": fl("hybridsearch on doc2"),
        "binary_rsf": fl("hybrid search with relative score fusion on binary vector"),
        "binary_rrf": fl("hybrid search with rank fusion on binary vector"),
    }
    s = sections(read("00041_mqvs_text_search_multiple_parts.reference"), lambda ln: not re.match(r"^\d+(\t|$)", ln))
    g["00041_multi_parts"] = {
        "source": "tests/queries/2_vector_search/00041_mqvs_text_search_multiple_parts.{sql,reference}",
        "parts": "ids 0..9 in part 0, ids 10..19 in part 1 (same 20 docs as 00040)",
        "text_2parts": [[int(x[0]), float(x[1])] for x in s["Text search result with 2 parts"]],
        "rsf_2parts_stale_per_part_fusion": [[int(x[0]), float(x[1])] for x in s["Hybrid search RSF result with 2 parts"]],
        "text_1part": [[int(x[0]), float(x[1])] for x in s["Text search result with 1 part after optimize final"]],
        "rsf_1part": [[int(x[0]), float(x[1])] for x in s["Hybrid search RSF result with 1 part after optimize final"]],
    }

    # 00003: PREWHERE filter on the 100-row FLAT table (filter bitmap path), top-20 ordered by (d, id)
    r = rows(read("00003_mqvs_distance_with_prewhere.reference"))
    g["00003_prewhere"] = {
        "source": "tests/queries/2_vector_search/00003_mqvs_distance_with_prewhere.{sh,reference}",
        "corpus": "row n = [n,n,n] for n in range(100)", "query": [1.0, 1.0, 1.0], "metric": "L2", "k": 20,
        "filter": "id < 10 or id > 60",
        "expect": [[int(x[0]), float(x[2])] for x in r],
    }

    # 00008: empty vectors (ids 10..29 are []), 430 rows, query [20]*3: first through the IVFFLAT index, then FLAT
    r = rows(read("00008_mqvs_empty_vector.reference"))
    g["00008_empty_vectors"] = {
        "source": "tests/queries/2_vector_search/00008_mqvs_empty_vector.{sh,reference} + helpers/00000_prepare_data_with_empty_vectors.sh",
        "corpus": "ids 0..9 = [n]*3, ids 10..29 = [] (empty), ids 30..429 = [n]*3", "query": [20.0] * 3, "metric": "L2", "k": 10,
        "expect_ivfflat": [[int(x[0]), float(x[2])] for x in r[:10]],
        "expect_flat": [[int(x[0]), float(x[2])] for x in r[10:20]],
    }

    # 00009 / 00011: brute force + PREWHERE on the 10030-row table with empty vectors (helper 2), top-100
    for name, key, flt, q in (("00009_mqvs_brute_force_search_prewhere_0", "00009_bruteforce_prewhere",
                               "id > 5000 or id in (9, 31, 999, 1)", [10020.1] * 3),
                              ("00011_mqvs_brute_force_search_where", "00011_bruteforce_prewhere_sparse",
                               "id < 50 or id in (51, 55, 99, 100, 9999)", [10020.0] * 3)):
        r = rows(read(name + ".reference"))
        g[key] = {
            "source": f"tests/queries/2_vector_search/{name}.{{sh,reference}} + helpers/00000_prepare_index_2.sh",
            "corpus": "ids 0..9 = [n]*3, ids 10..29 = [] (empty), ids 30..10029 = [n]*3; index_granularity=128",
            "query": q, "metric": "L2", "k": 100, "filter": flt,
            "expect": [[int(x[0]), float(x[2])] for x in r],
        }

    # 00016: lightweight delete of id = 2 on 2100 rows, top-10
    r = [x for x in rows(read("00016_mqvs_lightweight_delete_with_vector.reference")) if len(x) == 3]
    g["00016_lightweight_delete"] = {
        "source": "tests/queries/2_vector_search/00016_mqvs_lightweight_delete_with_vector.{sql,reference}",
        "corpus": "row n = [n,n,n] for n in range(2100); DELETE WHERE id = 2", "query": [0.1] * 3, "metric": "L2", "k": 10,
        "expect": [[int(x[0]), float(x[2])] for x in r],
    }

    # 00040 with lightweight delete: same 20 docs; DELETE id = 13 keeps the index statistics (the doc stays in the
    # segment, only the alive bitmap changes); hybrid RSF with an EMPTY text list (part without an FTS index)
    s = sections(read("00040_mqvs_hybrid_search_with_lwd.reference"), lambda ln: not re.match(r"^\d+(\t|$)", ln))
    pr = lambda h: [[int(x[0]), float(x[1])] for x in s[h]]
    g["00040_hybrid_with_lwd"] = {
        "source": "tests/queries/2_vector_search/00040_mqvs_hybrid_search_with_lwd.{sql,reference}",
        "docs": "the 20 docs of 00040_hybrid; vectors [n,n,n]", "query_text": "Ancient", "query_vector": [1.0, 1.0, 1.0],
        "deleted_id": 13,
        "text_before_lwd_top1": pr("text search before LWD"),
        "text_after_lwd_top1": pr("text search after LWD"),
        "text_no_index": pr("text search on part w/o tantivy index"),
        "rsf_no_text_index": pr("hybrid search rsf on part w/o tantivy index"),
        "text_after_lwd_top2": pr("text search on part with index after LWD"),
        "rsf_after_lwd": pr("hybrid search rsf on part with index after LWD"),
    }

    # 00029: MSTG on a part below min_bytes_to_build -> fallback to FLAT; cosine distances down at 1e-4 (1 - ip in fp32)
    r = [x for x in rows(read("00029_mqvs_fallback_to_flat.reference")) if len(x) == 2 and re.match(r"^\d+$", x[0])]
    g["00029_fallback_to_flat_cosine"] = {
        "source": "tests/queries/2_vector_search/00029_mqvs_fallback_to_flat.{sql,reference}",
        "corpus": "row n = [n, n+7, n+6, n+5, n+4, n+3, n+2, n+1] for n in range(1000)",
        "query": [8.0, 15, 14, 13, 12, 11, 10, 9], "metric": "COSINE", "k": 5,
        "expect": [[int(x[0]), float(x[1])] for x in r[:5]],
        "expect_after_reload": [[int(x[0]), float(x[1])] for x in r[5:10]],
    }

    with open(os.path.join(OUT, "reference_goldens.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", os.path.join(OUT, "reference_goldens.json"), "cases:", list(g))


if __name__ == "__main__":
    main()
