"""Independent (Python, regex) parser of a DBoW2 1.0.2 BRIEF vocabulary (`voc.yml.gz`).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Follows TemplatedVocabulary::load as restated
in SURVEY.md A.3: nodes are indexed by DBoW2 node id (0 = root), a node is appended to its
parent's child list in FILE order, descriptor strings are MSB-first (char c <-> bit 255-c) and
DBoW2 bit i <-> OpenCV byte i/8, bit i%8 (reference src/CBoWManager.h:102-106).
"""
import gzip
import re
import numpy as np

_NODE = re.compile(r'nodeId:(\d+),\s*parentId:(\d+),\s*weight:([-+0-9.eE]+),\s*descriptor:"([01]+)"')
_WORD = re.compile(r'wordId:(\d+),\s*nodeId:(\d+)')
_HEAD = re.compile(r'\b(k|L|scoringType|weightingType):\s*(\d+)')


def load(path):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        text = f.read()
    head = dict((k, int(v)) for k, v in _HEAD.findall(text[:400]))
    nodes = _NODE.findall(text)
    words = _WORD.findall(text)
    nn = len(nodes) + 1
    parent = np.zeros(nn, np.int32)
    weight = np.zeros(nn, np.float64)
    desc = np.zeros((nn, 32), np.uint8)
    children = [[] for _ in range(nn)]
    for nid, pid, w, d in nodes:
        nid = int(nid); pid = int(pid)
        parent[nid] = pid
        weight[nid] = float(w)
        children[pid].append(nid)
        bits = np.frombuffer(d.encode(), np.uint8) - ord("0")      # char c <-> bit 255-c
        assert bits.size == 256
        desc[nid] = np.packbits(bits[::-1], bitorder="little")     # bit i -> byte i/8, bit i%8
    node_word = np.full(nn, -1, np.int32)
    for wid, nid in words:
        node_word[int(nid)] = int(wid)
    child_ofs = np.zeros(nn + 1, np.int32)
    child_ofs[1:] = np.cumsum([len(c) for c in children])
    child_idx = np.array([c for cs in children for c in cs], np.int32)
    return {"k": head["k"], "L": head["L"], "scoring": head["scoringType"], "weighting": head["weightingType"],
            "nnodes": nn, "nwords": len(words), "parent": parent, "node_weight": weight,
            "node_desc": np.ascontiguousarray(desc), "node_word": node_word,
            "child_ofs": child_ofs, "child_idx": child_idx}
