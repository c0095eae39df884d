"""Loader for DBoW2 1.0.2 BRIEF vocabularies (`voc.yml.gz`), product side.

Mirrors CBoWManager::loadVocabularyFromConfigFile -> BriefVocabulary::load (reference
src/CBoWManager.h:59-66): produces the flattened arrays sfe_bow_create() takes -- nodes indexed by DBoW2
node id (0 = root), children in file order, descriptors repacked to OpenCV byte order (bit i of the DBoW2
bitset = byte i/8, bit i%8, CBoWManager.h:102-106; YAML strings are MSB-first).  The C++ adapter
(host/voc_loader.cpp) does the same for the C++ pipeline.  Line-streaming parser, no regex over the whole file.
"""
import gzip
import numpy as np


def _field(line, key):
    i = line.index(key) + len(key)
    j = i
    while j < len(line) and line[j] not in ",}\n ":
        j += 1
    return line[i:j]


def load_vocabulary(path):
    opener = gzip.open if str(path).endswith(".gz") else open
    parents, weights, ids, descs, words = [], [], [], [], []
    head = {}
    pending = None
    with opener(path, "rt") as f:
        for line in f:
            s = line.strip()
            if s.startswith("- { nodeId:"):
                pending = (int(_field(s, "nodeId:")), int(_field(s, "parentId:")), float(_field(s, "weight:")))
                if "descriptor:" not in s:
                    continue
            if "descriptor:" in s and pending is not None:
                q = s.index('"') + 1
                bits = s[q:s.index('"', q)]
                ids.append(pending[0]); parents.append(pending[1]); weights.append(pending[2]); descs.append(bits)
                pending = None
            elif s.startswith("- { wordId:"):
                words.append((int(_field(s, "wordId:")), int(_field(s, "nodeId:"))))
            else:
                for k in ("k", "L", "scoringType", "weightingType"):
                    if s.startswith(k + ":"):
                        head[k] = int(s.split(":")[1])
    nn = max(ids) + 1
    node_desc = np.zeros((nn, 32), np.uint8)
    node_weight = np.zeros(nn, np.float64)
    parent = np.zeros(nn, np.int32)
    counts = np.zeros(nn, np.int64)
    for nid, pid in zip(ids, parents):
        counts[pid] += 1
    child_ofs = np.zeros(nn + 1, np.int32)
    child_ofs[1:] = np.cumsum(counts)
    fill = child_ofs[:-1].copy()
    child_idx = np.zeros(len(ids), np.int32)
    for nid, pid, w, bits in zip(ids, parents, weights, descs):
        child_idx[fill[pid]] = nid; fill[pid] += 1
        parent[nid] = pid; node_weight[nid] = w
        b = np.frombuffer(bits.encode("ascii"), np.uint8) - 48            # MSB first: char c <-> bit 255-c
        node_desc[nid] = np.packbits(b[::-1], bitorder="little")
    node_word = np.full(nn, -1, np.int32)
    for wid, nid in words:
        node_word[nid] = wid
    if head.get("scoringType", 0) != 0 or head.get("weightingType", 0) != 0:
        raise ValueError("only L1 scoring + TF-IDF weighting (the shipped vocabulary) are supported")
    return {"k": head["k"], "L": head["L"], "scoring": head["scoringType"], "weighting": head["weightingType"],
            "nnodes": nn, "nwords": len(words), "parent": parent, "node_weight": node_weight, "node_desc": node_desc,
            "node_word": node_word, "child_ofs": child_ofs, "child_idx": child_idx}
