"""Flat proof buffer  <->  the reference's serde-JSON `Proof<F, H, EXT>`.

The reference stores proofs as pretty serde-JSON of boojum's `Proof` wrapped in an externally tagged enum
(`{"MainVM": {...}}`, /root/reference/src/data_source/local_file_data_source.rs:51-55, type aliases
circuit_definitions/src/circuit_definitions/base_layer/mod.rs:517-529); the golden files under test_proofs/ are that format.
libzkgpu returns one flat u64 buffer whose field order is the JSON's field order after a 32-word header
(csrc/prover.cu "assemble the proof"):

  header[32]: magic, log_n, log_lde, cap, n_queries, n_fri_oracles, W, S2, Q, S, n_at_z, n_at_z_omega, n_at_0, n_public_inputs,
              n_final, pow_bits, fri_schedule[16]
  public_inputs | witness cap | stage-2 cap | quotient cap | final_fri_monomials (c0.., c1..) | values_at_z (c0,c1)..
  | values_at_z_omega | values_at_0 | fri caps (base, intermediates) |
  per query: witness / stage_2 / quotient / setup {leaf_elements, path} then per FRI oracle {leaf_elements, path} | pow_challenge

`proof_to_dict` / `proof_from_dict` convert both ways, so a proof produced here can be written where the reference's
`LocalFileDataSource` expects it, and a reference proof can be loaded into the flat layout.
"""
import json

import numpy as np

PROOF_MAGIC = 0x5A4B50524F4F4631


def _ext(c0, c1):
    return {"coeffs": [int(c0), int(c1)], "_marker": None}


class _Reader:
    def __init__(self, a):
        self.a, self.i = a, 0

    def take(self, n):
        v = self.a[self.i:self.i + n]
        assert len(v) == n, "proof buffer too short"
        self.i += n
        return v


def parse_header(flat):
    h = [int(x) for x in flat[:32]]
    if h[0] != PROOF_MAGIC:
        raise ValueError("not a zkgpu proof buffer")
    keys = ["log_n", "log_lde", "cap", "n_queries", "n_fri", "W", "S2", "Q", "S", "n_at_z", "n_at_zw", "n_at_0", "n_pi", "n_final", "pow_bits"]
    hd = dict(zip(keys, h[1:16]))
    hd["schedule"] = h[16:16 + hd["n_fri"]]
    return hd


def _fri_shapes(hd):
    """(leaves, cap entries, path length) per FRI oracle, as csrc/host.cu make_shape derives them"""
    out, log_dom = [], hd["log_n"] + hd["log_lde"]
    for s in hd["schedule"]:
        leaves = 1 << (log_dom - s)
        cap = min(hd["cap"], leaves)
        out.append((leaves, cap, (leaves // cap).bit_length() - 1))
        log_dom -= s
    return out


def proof_to_dict(flat, variant=None, security_level=None):
    """flat u64 buffer -> dict with the reference's JSON structure ({variant: {...}} when `variant` is given)."""
    flat = np.asarray(flat, dtype=np.uint64)
    hd = parse_header(flat)
    r = _Reader(flat)
    r.take(32)
    cap, depth = hd["cap"], hd["log_n"] + hd["log_lde"] - (hd["cap"].bit_length() - 1)
    fri = _fri_shapes(hd)

    def digests(n):
        return [[int(x) for x in r.take(4)] for _ in range(n)]

    def exts(n):
        v = r.take(2 * n)
        return [_ext(v[2 * i], v[2 * i + 1]) for i in range(n)]

    p = {}
    p["proof_config"] = {"fri_lde_factor": 1 << hd["log_lde"], "merkle_tree_cap_size": cap, "fri_folding_schedule": None,
                         "security_level": security_level if security_level is not None else hd["n_queries"] * hd["log_lde"],
                         "pow_bits": hd["pow_bits"]}
    p["public_inputs"] = [int(x) for x in r.take(hd["n_pi"])]
    p["witness_oracle_cap"] = digests(cap)
    p["stage_2_oracle_cap"] = digests(cap)
    p["quotient_oracle_cap"] = digests(cap)
    p["final_fri_monomials"] = [[int(x) for x in r.take(hd["n_final"])], [int(x) for x in r.take(hd["n_final"])]]
    p["values_at_z"] = exts(hd["n_at_z"])
    p["values_at_z_omega"] = exts(hd["n_at_zw"])
    p["values_at_0"] = exts(hd["n_at_0"])
    caps = [digests(c) for (_, c, _) in fri]
    p["fri_base_oracle_cap"] = caps[0]
    p["fri_intermediate_oracles_caps"] = caps[1:]
    queries = []
    for _ in range(hd["n_queries"]):
        q = {}
        for name, width in (("witness_query", hd["W"]), ("stage_2_query", hd["S2"]), ("quotient_query", hd["Q"]), ("setup_query", hd["S"])):
            q[name] = {"leaf_elements": [int(x) for x in r.take(width)], "proof": digests(depth)}
        q["fri_queries"] = [{"leaf_elements": [int(x) for x in r.take(2 << s)], "proof": digests(d)} for s, (_, _, d) in zip(hd["schedule"], fri)]
        queries.append(q)
    p["queries_per_fri_repetition"] = queries
    p["pow_challenge"] = int(r.take(1)[0])
    p["_marker"] = None
    assert r.i == flat.size, "trailing words in proof buffer"
    return {variant: p} if variant else p


def proof_from_dict(d, log_n=None):
    """reference JSON structure (with or without the enum tag) -> (flat u64 buffer, variant name)."""
    variant = None
    if "proof_config" not in d:
        (variant, d), = d.items()
    pc = d["proof_config"]
    log_lde = pc["fri_lde_factor"].bit_length() - 1
    cap = pc["merkle_tree_cap_size"]
    q0 = d["queries_per_fri_repetition"][0]
    depth = len(q0["witness_query"]["proof"])
    if log_n is None:
        log_n = depth + (cap.bit_length() - 1) - log_lde
    schedule = [(len(f["leaf_elements"]) // 2).bit_length() - 1 for f in q0["fri_queries"]]
    W, S2, Q, S = (len(q0[k]["leaf_elements"]) for k in ("witness_query", "stage_2_query", "quotient_query", "setup_query"))
    hdr = [PROOF_MAGIC, log_n, log_lde, cap, len(d["queries_per_fri_repetition"]), len(schedule), W, S2, Q, S, len(d["values_at_z"]),
           len(d["values_at_z_omega"]), len(d["values_at_0"]), len(d["public_inputs"]), len(d["final_fri_monomials"][0]), pc["pow_bits"]]
    hdr += schedule + [0] * (32 - len(hdr) - len(schedule))
    out = list(hdr) + list(d["public_inputs"])
    for key in ("witness_oracle_cap", "stage_2_oracle_cap", "quotient_oracle_cap"):
        for dg in d[key]:
            out += dg
    out += d["final_fri_monomials"][0] + d["final_fri_monomials"][1]
    for key in ("values_at_z", "values_at_z_omega", "values_at_0"):
        for e in d[key]:
            out += e["coeffs"]
    for c in [d["fri_base_oracle_cap"]] + list(d["fri_intermediate_oracles_caps"]):
        for dg in c:
            out += dg
    for q in d["queries_per_fri_repetition"]:
        for key in ("witness_query", "stage_2_query", "quotient_query", "setup_query"):
            out += q[key]["leaf_elements"]
            for dg in q[key]["proof"]:
                out += dg
        for f in q["fri_queries"]:
            out += f["leaf_elements"]
            for dg in f["proof"]:
                out += dg
    out.append(d["pow_challenge"])
    return np.array(out, dtype=np.uint64), variant


def save_proof_json(path, flat, variant, security_level=None):
    """writes the proof where the reference's LocalFileDataSource would read it (pretty JSON, externally tagged)"""
    with open(path, "w") as f:
        json.dump(proof_to_dict(flat, variant, security_level), f, indent=2)


def load_proof_json(path):
    return proof_from_dict(json.load(open(path)))
