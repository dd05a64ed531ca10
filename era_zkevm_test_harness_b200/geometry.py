"""Circuit geometry, gate tables and proof configs -- the host-side mirror of the reference's
`circuit_definitions` shapes as ctypes structs of include/zkgpu.h.

Reference:
  * geometry per circuit: circuit_definitions/src/circuit_definitions/base_layer/*.rs (`geometry()`, `lookup_parameters()`,
    `configure_builder()`), recursion_layer/{circuit_def,scheduler}.rs
  * proof configs: circuit_definitions/src/lib.rs:13-57 (`base_layer_proof_config()` ...)
  * VK JSON (`VerificationKey.fixed_parameters`): setup/base_layer/vk_N.json, setup/recursion_layer/vk_*.json
"""
import ctypes
import json
import math

MAX_GATES = 24
MAX_PI = 8
MAX_FRI = 16

# gate kinds (include/zkgpu.h)
(GATE_NOP, GATE_CONSTANTS_ALLOCATOR, GATE_FMA, GATE_REDUCTION4, GATE_SELECTION, GATE_PARALLEL_SELECTION4, GATE_ZERO_CHECK,
 GATE_UINTX_ADD, GATE_DOT_PRODUCT4, GATE_U8X4_FMA, GATE_POSEIDON2_FLATTENED, GATE_FMA_EXT, GATE_U32_TRI_ADD_CARRY,
 GATE_BOUNDED_BOOLEAN, GATE_MATMUL12_EXTERNAL, GATE_MATMUL12_INNER, GATE_NONLINEARITY7, GATE_CONDITIONAL_SWAP4,
 GATE_ZERO_CHECK_WITNESS, GATE_BOOLEAN_ALL) = range(20)

GATE_NAMES = {
    "ConstantsAllocator": GATE_CONSTANTS_ALLOCATOR, "FmaBaseNoConst": GATE_FMA, "Reduction4": GATE_REDUCTION4,
    "Selection": GATE_SELECTION, "ParallelSelection4": GATE_PARALLEL_SELECTION4, "ZeroCheck": GATE_ZERO_CHECK,
    "UIntXAdd": GATE_UINTX_ADD, "DotProduct4": GATE_DOT_PRODUCT4, "U8x4FMA": GATE_U8X4_FMA,
    "Poseidon2Flattened": GATE_POSEIDON2_FLATTENED, "FmaExt": GATE_FMA_EXT, "PublicInput": GATE_NOP, "Nop": GATE_NOP,
    "U32TriAddCarryAsChunk": GATE_U32_TRI_ADD_CARRY,  # StorageApplication only
    # compression circuits (aux_layer/compression_modes/mode_{1..4}.rs)
    "BoundedBoolean": GATE_BOUNDED_BOOLEAN, "MatMul12External": GATE_MATMUL12_EXTERNAL, "MatMul12Inner": GATE_MATMUL12_INNER,
    "Nonlinearity7": GATE_NONLINEARITY7, "ConditionalSwap4": GATE_CONDITIONAL_SWAP4, "ZeroCheckWitness": GATE_ZERO_CHECK_WITNESS,
    "BooleanAllColumns": GATE_BOOLEAN_ALL,  # EIP-4844: BooleanConstraintGate on general-purpose columns
}

# gate_idx -> gate name per verification key, derived in SURVEY.md section 8a by matching each VK's
# (num_constants, degree) list against the configure_builder order of the circuit
BASE_LAYER_GATE_ORDER = {
    1: ["ConstantsAllocator", "U8x4FMA", "Poseidon2Flattened", "DotProduct4", "ZeroCheck", "FmaBaseNoConst", "UIntXAdd", "Selection",
        "ParallelSelection4", "PublicInput", "Reduction4"],
    7: ["ConstantsAllocator", "U8x4FMA", "ZeroCheck", "FmaBaseNoConst", "UIntXAdd", "DotProduct4", "Selection", "ParallelSelection4",
        "PublicInput", "Reduction4"],
    10: ["ConstantsAllocator", "ZeroCheck", "FmaBaseNoConst", "UIntXAdd", "U32TriAddCarryAsChunk", "Selection", "ParallelSelection4",
         "PublicInput", "Reduction4"],
}
for _t in (2, 4, 8, 9, 11, 12):
    BASE_LAYER_GATE_ORDER[_t] = ["ConstantsAllocator", "Poseidon2Flattened", "ZeroCheck", "FmaBaseNoConst", "UIntXAdd", "Selection",
                                 "ParallelSelection4", "PublicInput", "Reduction4"]
for _t in (3, 6):
    BASE_LAYER_GATE_ORDER[_t] = ["ConstantsAllocator", "FmaBaseNoConst", "Reduction4", "Selection", "ParallelSelection4", "PublicInput",
                                 "UIntXAdd", "ZeroCheck"]
for _t in (5, 13):
    BASE_LAYER_GATE_ORDER[_t] = ["ConstantsAllocator", "ZeroCheck", "FmaBaseNoConst", "UIntXAdd", "Selection", "ParallelSelection4",
                                 "PublicInput", "Reduction4"]
# compression_{N}_vk.json: gate_idx follows the configure_builder order of mode_N.rs; (num_constants, degree) per gate_idx
# match the VKs: Poseidon2Flattened (0, 7), MatMul12 (0, 1) x 2, Nonlinearity7 (1, 7), ConditionalSwap4 (0, 2), ...
_COMPRESSION_TAIL = ["FmaBaseNoConst", "FmaExt", "Selection", "ParallelSelection4", "ConditionalSwap4", "PublicInput", "Reduction4"]
COMPRESSION_GATE_ORDER = {
    # mode 1: BooleanConstraintGate sits in its own specialised column (has_boolean_col), ZeroCheck keeps its inverse in a witness column
    1: ["ConstantsAllocator", "Poseidon2Flattened", "ZeroCheckWitness"] + _COMPRESSION_TAIL,
    2: ["ConstantsAllocator", "BoundedBoolean", "Poseidon2Flattened", "ZeroCheckWitness"] + _COMPRESSION_TAIL,
    3: ["ConstantsAllocator", "BoundedBoolean", "Poseidon2Flattened", "ZeroCheckWitness"] + _COMPRESSION_TAIL,
    # mode 4: no plain witness columns; the round function is built from matrix-multiplication and x^7 gates (48 columns < 130)
    4: ["ConstantsAllocator", "BoundedBoolean", "MatMul12External", "MatMul12Inner", "Nonlinearity7", "ZeroCheck"] + _COMPRESSION_TAIL,
}
RECURSION_GATE_ORDER = ["ConstantsAllocator", "Poseidon2Flattened", "ZeroCheck", "FmaBaseNoConst", "FmaExt", "UIntXAdd", "Selection",
                        "ParallelSelection4", "PublicInput", "Reduction4"]

# EIP-4844 circuit (circuit_definitions/src/circuit_definitions/eip4844/mod.rs:43-112, VK setup/aux_layer/eip4844_vk.json):
# 60 copy columns, 8 constant columns, lookup 3x20, NO specialised boolean column -- BooleanConstraintGate sits on the
# general-purpose columns (one instance per column); (num_constants, degree) per gate_idx match the VK:
# (8,1) (0,0) (2,3) (4,2) (0,2) (1,2) (0,2) (0,2).  UIntXAddGate<32>/<16> collapse into one entry like everywhere else.
EIP4844_GATE_ORDER = ["ConstantsAllocator", "PublicInput", "FmaBaseNoConst", "Reduction4", "BooleanAllColumns", "UIntXAdd", "Selection",
                      "DotProduct4"]

BASE_LAYER_CIRCUIT_NAMES = {
    1: "MainVM", 2: "CodeDecommittmentsSorter", 3: "CodeDecommitter", 4: "LogDemuxer", 5: "KeccakRoundFunction", 6: "Sha256RoundFunction",
    7: "ECRecover", 8: "RAMPermutation", 9: "StorageSorter", 10: "StorageApplication", 11: "EventsSorter", 12: "L1MessagesSorter",
    13: "L1MessagesHasher",
}


class Gate(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("n_consts", ctypes.c_uint32), ("path_len", ctypes.c_uint32), ("path_bits", ctypes.c_uint32)]


class Geometry(ctypes.Structure):
    _fields_ = [("log_n", ctypes.c_uint32), ("n_copy", ctypes.c_uint32), ("n_witness_plain", ctypes.c_uint32), ("n_const_cols", ctypes.c_uint32),
                ("lookup_width", ctypes.c_uint32), ("lookup_reps", ctypes.c_uint32), ("table_id_col", ctypes.c_uint32),
                ("has_boolean_col", ctypes.c_uint32), ("quotient_degree", ctypes.c_uint32), ("table_len", ctypes.c_uint32),
                ("n_public_inputs", ctypes.c_uint32), ("pi_col", ctypes.c_uint32 * MAX_PI), ("pi_row", ctypes.c_uint32 * MAX_PI),
                ("n_gates", ctypes.c_uint32), ("gates", Gate * MAX_GATES)]

    # derived counts (same formulas as csrc/host.cu make_shape)
    @property
    def n_perm(self):
        return self.n_copy + (1 if self.has_boolean_col else 0) + self.lookup_width * self.lookup_reps

    @property
    def n_witness(self):  # copy-permuted columns, plain witness columns, lookup multiplicities
        return self.n_perm + self.n_witness_plain + (1 if self.lookup_reps else 0)

    @property
    def n_setup(self):
        return self.n_perm + self.n_const_cols + (self.lookup_width + 1 if self.lookup_reps else 0)

    @property
    def n_stage2(self):
        c = (self.n_perm + self.quotient_degree - 1) // self.quotient_degree
        return 2 * (c + self.lookup_reps + (1 if self.lookup_reps else 0))

    @property
    def n_quotient(self):
        return 2 * self.quotient_degree

    def scaled(self, log_n, table_len=None):
        """same circuit shape on a shorter trace (parity-test sizes); public-input rows and table length are clamped"""
        g = Geometry.from_buffer_copy(bytes(self))
        g.log_n = log_n
        n = 1 << log_n
        for i in range(g.n_public_inputs):
            g.pi_row[i] = self.pi_row[i] % n
        g.table_len = min(self.table_len, n // 2) if table_len is None else table_len
        return g


class ProofConfig(ctypes.Structure):
    _fields_ = [("log_lde", ctypes.c_uint32), ("cap_size", ctypes.c_uint32), ("n_queries", ctypes.c_uint32), ("pow_bits", ctypes.c_uint32),
                ("n_fri_oracles", ctypes.c_uint32), ("fri_schedule", ctypes.c_uint32 * MAX_FRI)]


def fri_schedule(log_n, log_lde, cap_size):
    """Folding schedule as the golden proofs show it: fold by 8 while (a) the new oracle still has at least cap_size leaves
    and (b) the polynomial still has degree bits left; the last oracle folds by less, and what is left is the final
    polynomial.  2^20, lde 2, cap 16 -> [3,3,3,3,3,2] + 8 final monomials (base/recursion goldens); 2^20, cap 32 ->
    [3,3,3,3,3,1] (proof.json); compression modes: 2^16 lde 32 -> [3,3,3,3,3,1], 2^13 lde 512 -> [3,3,3,3,1],
    2^12 lde 1024 -> [3,3,3,3], 2^15 lde 2048 cap 256 -> [3,3,3,3,3], all with ONE final monomial
    (compression_{1..4}_proof.json)."""
    log_dom, deg_bits, log_cap = log_n + log_lde, log_n, int(math.log2(cap_size))
    sched = []
    while True:
        s = min(3, deg_bits, log_dom - log_cap)
        if s <= 0:
            break
        sched.append(s)
        deg_bits -= s
        log_dom -= s
    return sched or [1]


def make_proof_config(log_n, fri_lde_factor=2, merkle_tree_cap_size=16, security_level=100, pow_bits=0, schedule=None):
    """ProofConfig{fri_lde_factor, merkle_tree_cap_size, fri_folding_schedule: None, security_level, pow_bits}
    (circuit_definitions/src/lib.rs:29-37) -> derived query count and folding schedule."""
    log_lde = int(math.log2(fri_lde_factor))
    assert 1 << log_lde == fri_lde_factor
    cfg = ProofConfig()
    cfg.log_lde = log_lde
    cfg.cap_size = merkle_tree_cap_size
    cfg.n_queries = -(-security_level // log_lde)
    cfg.pow_bits = pow_bits
    sched = fri_schedule(log_n, log_lde, merkle_tree_cap_size) if schedule is None else list(schedule)
    cfg.n_fri_oracles = len(sched)
    for i, s in enumerate(sched):
        cfg.fri_schedule[i] = s
    return cfg


def recursion_layer_proof_config(log_n=20):
    """circuit_definitions/src/lib.rs:39-47 `recursion_layer_proof_config()`: lde 2, cap 16, security 100, no PoW."""
    return make_proof_config(log_n, fri_lde_factor=2, merkle_tree_cap_size=16, security_level=100, pow_bits=0)


def eip4844_proof_config(log_n=20):
    """circuit_definitions/src/lib.rs:49-57 `eip4844_proof_config()`: the base-layer constants."""
    return base_layer_proof_config(log_n)


def base_layer_proof_config(log_n=20):
    """circuit_definitions/src/lib.rs:29-37: lde 2, cap 16, security 100, pow 0"""
    return make_proof_config(log_n, 2, 16, 100, 0)


recursion_layer_proof_config = base_layer_proof_config  # lib.rs:39-47 -- same constants

# aux_layer/compression_modes/mode_{1..4}.rs: (trace length, fri_lde_factor, merkle_tree_cap_size); security = L1_SECURITY_BITS = 80
COMPRESSION_MODES = {1: (16, 32, 16), 2: (13, 512, 16), 3: (12, 1024, 16), 4: (15, 2048, 256)}


def compression_layer_proof_config(mode, log_n=None):
    """ProofConfig of CompressionMode{mode} (mode_N.rs `proof_config_for_compression_step`): queries 16 / 9 / 8 / 8.
    The circuits themselves (plain witness columns, BoundedBoolean / ConditionalSwap / matrix-multiplication gates) come from
    the compression VKs: compression_geometries_from_fixture."""
    ln, lde, cap = COMPRESSION_MODES[mode]
    return make_proof_config(ln if log_n is None else log_n, lde, cap, security_level=80)


def _walk_selector_tree(node, path, out):
    """selectors_placement is {"Fork": {"left": .., "right": ..}} / {"GateOnly": {"gate_idx", "num_constants", "degree_of_gate", ..}} /
    "Empty"; left = constant bit 0, right = bit 1 (MainVM: Poseidon2Flattened sits at path [0], ConstantsAllocator at [1,1,1])"""
    if node == "Empty" or node is None:
        return
    if "Fork" in node:
        f = node["Fork"]
        _walk_selector_tree(f["left"], path + [0], out)
        _walk_selector_tree(f["right"], path + [1], out)
        return
    g = node.get("GateOnly") or node.get("Gate") or node
    out.append((g["gate_idx"], g["num_constants"], g.get("degree_of_gate", g.get("degree")), list(path)))


def geometry_from_vk(vk, gate_order, has_boolean_col=1):
    """vk: the inner dict of a VK JSON file ({"fixed_parameters": .., "setup_merkle_tree_cap": ..});
    gate_order: gate names by gate_idx (BASE_LAYER_GATE_ORDER[type] / RECURSION_GATE_ORDER)."""
    fp = vk["fixed_parameters"]
    par = fp["parameters"]
    g = Geometry()
    g.log_n = int(math.log2(fp["domain_size"]))
    g.n_copy = par["num_columns_under_copy_permutation"]
    g.n_witness_plain = par["num_witness_columns"]  # compression modes 1-3 only
    lp = fp["lookup_parameters"]
    if lp == "NoLookup":
        g.lookup_width = g.lookup_reps = 0
    else:
        (kind, body), = lp.items()
        assert kind == "UseSpecializedColumnsWithTableIdAsConstant", kind
        g.lookup_width, g.lookup_reps = body["width"], body["num_repetitions"]
    g.n_const_cols = par["num_constant_columns"] + fp["extra_constant_polys_for_selectors"] + (1 if g.lookup_reps else 0)
    g.table_id_col = fp["table_ids_column_idxes"][0] if g.lookup_reps else 0
    # every base/recursion circuit (and compression mode 1) places BooleanConstraintGate in its own specialised column;
    # compression modes 2-4 use BoundedBooleanConstraintGate on general-purpose columns instead
    g.has_boolean_col = has_boolean_col
    g.quotient_degree = fp["quotient_degree"]
    g.table_len = fp["total_tables_len"]
    pis = fp["public_inputs_locations"]
    g.n_public_inputs = len(pis)
    for i, (col, row) in enumerate(pis):
        g.pi_col[i], g.pi_row[i] = col, row
    gates = []
    _walk_selector_tree(fp["selectors_placement"], [], gates)
    gates.sort()
    g.n_gates = len(gates)
    for i, (idx, n_consts, _deg, path) in enumerate(gates):
        assert idx == i, "gate indexes are expected to be dense"
        g.gates[i].kind = GATE_NAMES[gate_order[idx]]
        g.gates[i].n_consts = n_consts
        g.gates[i].path_len = len(path)
        g.gates[i].path_bits = sum(b << k for k, b in enumerate(path))
    return g


def load_vk_json(path):
    d = json.load(open(path))
    if "fixed_parameters" not in d:
        (name, d), = d.items()
    else:
        name = None
    return name, d


def mainvm_like_geometry(log_n=20):
    """MainVM shape without reading the reference: 130 copy cols, lookup 3x8, 7+1 constant
    columns, the 11 gates of vk_1.json with their selector paths, quotient degree 8, 4 public inputs."""
    g = Geometry()
    g.log_n = log_n
    g.n_copy, g.n_const_cols = 130, 8
    g.lookup_width, g.lookup_reps, g.table_id_col = 3, 8, 7
    g.has_boolean_col, g.quotient_degree = 1, 8
    g.table_len = min(68756, (1 << log_n) // 2)
    g.n_public_inputs = 4
    for i in range(4):
        g.pi_col[i], g.pi_row[i] = i, 1033357 % (1 << log_n)
    spec = [("ConstantsAllocator", 4, "111"), ("U8x4FMA", 0, "100100"), ("Poseidon2Flattened", 0, "0"), ("DotProduct4", 0, "100010"),
            ("ZeroCheck", 0, "100011"), ("FmaBaseNoConst", 2, "10000"), ("UIntXAdd", 1, "101"), ("Selection", 0, "100110"),
            ("ParallelSelection4", 0, "100101"), ("PublicInput", 0, "100111"), ("Reduction4", 4, "110")]
    g.n_gates = len(spec)
    for i, (name, nc, path) in enumerate(spec):
        g.gates[i].kind = GATE_NAMES[name]
        g.gates[i].n_consts = nc
        g.gates[i].path_len = len(path)
        g.gates[i].path_bits = sum(int(b) << k for k, b in enumerate(path))
    return g


def small_test_geometry(log_n=8, n_copy=16, lookup=True):
    """a tiny circuit for fast CPU tests: FMA / Reduction / Selection / ZeroCheck / UIntXAdd / ConstantsAllocator"""
    g = Geometry()
    g.log_n = log_n
    g.n_copy = n_copy
    g.n_const_cols = 8 if lookup else 7
    if lookup:
        g.lookup_width, g.lookup_reps, g.table_id_col = 3, 2, 7
        g.table_len = (1 << log_n) // 4
    g.has_boolean_col, g.quotient_degree = 1, 8
    g.n_public_inputs = 2
    for i in range(2):
        g.pi_col[i], g.pi_row[i] = i, (1 << log_n) - 3
    spec = [("ConstantsAllocator", 4, "111"), ("FmaBaseNoConst", 2, "10000"), ("Reduction4", 4, "110"), ("Selection", 0, "100110"),
            ("ZeroCheck", 0, "100011"), ("UIntXAdd", 1, "101"), ("PublicInput", 0, "100111"), ("DotProduct4", 0, "100010")]
    g.n_gates = len(spec)
    for i, (name, nc, path) in enumerate(spec):
        g.gates[i].kind = GATE_NAMES[name]
        g.gates[i].n_consts = nc
        g.gates[i].path_len = len(path)
        g.gates[i].path_bits = sum(int(b) << k for k, b in enumerate(path))
    return g


def circuit_geometries_from_fixture(fixture):
    """fixture: the dict of tests/golden/vk_shapes.json (tools/make_vk_fixtures.py).  Yields (key, Geometry, entry) for
    the 13 base-layer circuits and the recursion-layer scheduler / leaf / node circuits of the reference."""
    for t, entry in sorted(fixture["base"].items(), key=lambda kv: int(kv[0])):
        yield f"base_{t}_{entry['variant']}", geometry_from_vk(entry, BASE_LAYER_GATE_ORDER[int(t)]), entry
    for key, entry in fixture["recursion"].items():
        yield f"recursion_{key}", geometry_from_vk(entry, RECURSION_GATE_ORDER), entry
    if "eip4844" in fixture.get("aux", {}):
        entry = fixture["aux"]["eip4844"]
        yield "aux_eip4844", geometry_from_vk(entry, EIP4844_GATE_ORDER, has_boolean_col=0), entry


def compression_geometries_from_fixture(fixture):
    """(key, Geometry, ProofConfig, entry) for the compression-layer circuits of the fixture: modes 1-4 as the one-shot
    compression_{N}_vk.json describe them, plus the wrapper-facing variants (same circuits, their own proof configs)."""
    for key, entry in fixture.get("compression", {}).items():
        mode = entry["mode"]
        geo = geometry_from_vk(entry, COMPRESSION_GATE_ORDER[mode], has_boolean_col=1 if mode == 1 else 0)
        fp = entry["fixed_parameters"]
        sec = entry["proof_shapes"][0]["proof_config"]["security_level"] if entry.get("proof_shapes") else 80
        cfg = make_proof_config(geo.log_n, fp["fri_lde_factor"], fp["cap_size"], security_level=sec)
        yield f"compression_{key}", geo, cfg, entry
