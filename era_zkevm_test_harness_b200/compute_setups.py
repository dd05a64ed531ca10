"""Host-side mirror of the reference's `src/compute_setups.rs` drivers over the C ABI.

  generate_circuit_setup_data   <-> compute_setups.rs:316-401  (one circuit type -> CircuitSetupData)
  generate_base_layer_vks       <-> compute_setups.rs:412-436  (loop over all basic circuits, store the VKs)
  generate_recursive_layer_vks  <-> compute_setups.rs:439-586  (leaf types, node, scheduler)
  generate_recursive_layer_vks_and_proofs <-> the same function's self-check: every leaf and the node circuit is PROVEN once with a
                                    placeholder witness right after its setup and the proof verified against the fresh VK
                                    (compute_setups.rs:479-496, :521-538), so a setup that cannot prove never reaches disk

and of the file names the reference's `LocalFileDataSource` uses for them
(src/data_source/local_file_data_source.rs:59-64, :95-113, :237-296): `{root}/base_layer/vk_{type}.json`,
`{root}/recursion_layer/vk_{type}.json`, `{root}/recursion_layer/vk_node.json`; pretty serde-JSON of
`{"<Variant>": {"fixed_parameters": .., "setup_merkle_tree_cap": [[u64;4]; cap]}}`.

The reference synthesises each circuit with an empty witness to obtain its setup columns (constants, sigmas, tables);
Rust synthesis cannot run in this image, so the setup columns come from `trace_source(geometry)` -- by default the synthetic
satisfying trace generator -- while the VK's `fixed_parameters` are carried over from the circuit's description.  The GPU
work is the reference's: LDE + Poseidon2 Merkle commitment of the setup columns (`get_full_setup`).
"""
import json
import os

import numpy as np

from . import geometry as G
from . import prover_utils as PU


class CircuitSetupData:
    """compute_setups.rs:303-312, collapsed like prover_utils.SetupData: device-resident setup + VK (+ optional variable maps)"""

    def __init__(self, key, variant, fixed_parameters, setup: PU.SetupData):
        self.key, self.variant, self.fixed_parameters, self.setup = key, variant, fixed_parameters, setup

    @property
    def vk(self):
        return {self.variant: {"fixed_parameters": self.fixed_parameters,
                               "setup_merkle_tree_cap": [[int(x) for x in row] for row in np.asarray(self.setup.vk_cap)]}}


def _default_trace_source(geo):
    return PU.synth_trace(geo, seed=0x5E7)[1]


def generate_circuit_setup_data(ctx, key, geo, entry, cfg=None, trace_source=_default_trace_source):
    cfg = cfg or G.base_layer_proof_config(geo.log_n)
    setup_cols = trace_source(geo)
    sd = PU.create_setup_data(ctx, geo, cfg, setup_cols)
    fp = dict(entry["fixed_parameters"])
    # a scaled-down trace (parity-test sizes) changes the size-dependent fields; at the reference's 2^20 these are no-ops
    fp["domain_size"] = 1 << geo.log_n
    fp["total_tables_len"] = int(geo.table_len)
    fp["public_inputs_locations"] = [[int(geo.pi_col[i]), int(geo.pi_row[i])] for i in range(geo.n_public_inputs)]
    return CircuitSetupData(key, entry["variant"], fp, sd)


def _synthetic_root(root):
    from .block import synthetic_root
    return synthetic_root(root)


def _write(path, obj):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(obj, f, indent=2)


def generate_base_layer_vks(ctx, root, fixture, log_n=None, trace_source=_default_trace_source):
    """-> {circuit_type: path}; one VK file per basic circuit type, named as the reference names them"""
    out = {}
    for t, entry in sorted(fixture["base"].items(), key=lambda kv: int(kv[0])):
        geo = G.geometry_from_vk(entry, G.BASE_LAYER_GATE_ORDER[int(t)])
        if log_n is not None and log_n != geo.log_n:
            geo = geo.scaled(log_n)
        data = generate_circuit_setup_data(ctx, f"base_{t}", geo, entry, trace_source=trace_source)
        path = os.path.join(_synthetic_root(root), "base_layer", f"vk_{t}.json")
        _write(path, data.vk)
        data.setup.close()
        out[int(t)] = path
    return out


def generate_recursive_layer_vks(ctx, root, fixture, log_n=None, trace_source=_default_trace_source):
    names = {"scheduler": "vk_1.json", "leaf_3": "vk_3.json", "node": "vk_node.json"}
    out = {}
    for key, entry in fixture["recursion"].items():
        geo = G.geometry_from_vk(entry, G.RECURSION_GATE_ORDER)
        if log_n is not None and log_n != geo.log_n:
            geo = geo.scaled(log_n)
        data = generate_circuit_setup_data(ctx, f"recursion_{key}", geo, entry, trace_source=trace_source)
        path = os.path.join(_synthetic_root(root), "recursion_layer", names[key])
        _write(path, data.vk)
        data.setup.close()
        out[key] = path
    return out


def generate_recursive_layer_vks_and_proofs(ctx, root, fixture, log_n=None, trace_source=_default_trace_source, witness_source=None):
    """compute_setups.rs:439-586: set up every recursion-layer circuit type, prove it once with a placeholder witness, verify the
    proof against the VK just computed, then store VK (and proof).  The reference's placeholder is the circuit synthesised over
    default (empty-queue) inputs; here it is a synthetic satisfying trace of the circuit's geometry (`witness_source(geo)`).
    -> {key: {"vk": path, "proof": path, "prove_seconds": s}}; raises if a proof does not verify (the reference asserts)."""
    import time
    from . import proof_format
    names = {"scheduler": ("vk_1.json", "scheduler_proof.json", "SchedulerCircuit"),
             "leaf_3": ("vk_3.json", "leaf_layer_proof_3_placeholder.json", "LeafLayerCircuit"),
             "node": ("vk_node.json", "node_layer_proof_placeholder.json", "NodeLayerCircuit")}
    if witness_source is None:
        witness_source = lambda geo: PU.synth_trace(geo, seed=0x5E7)[0]   # noqa: E731  (the witness that goes with the default setup)
    out = {}
    for key, entry in fixture["recursion"].items():
        geo = G.geometry_from_vk(entry, G.RECURSION_GATE_ORDER)
        if log_n is not None and log_n != geo.log_n:
            geo = geo.scaled(log_n)
        cfg = G.recursion_layer_proof_config(geo.log_n) if hasattr(G, "recursion_layer_proof_config") else G.base_layer_proof_config(geo.log_n)
        data = generate_circuit_setup_data(ctx, f"recursion_{key}", geo, entry, cfg=cfg, trace_source=trace_source)
        t0 = time.time()
        proof = PU.prove_circuit(ctx, data.setup, witness_source(geo))
        dt = time.time() - t0
        ok, msg = PU.verify_proof(geo, cfg, data.setup.vk_cap, proof)
        if not ok:
            data.setup.close()
            raise RuntimeError(f"compute_setups: the placeholder proof of {key} does not verify against its fresh VK: {msg}")
        vk_name, proof_name, variant = names[key]
        vk_path = os.path.join(_synthetic_root(root), "recursion_layer", vk_name)
        proof_path = os.path.join(_synthetic_root(root), "recursion_layer", proof_name)
        _write(vk_path, data.vk)
        proof_format.save_proof_json(proof_path, proof, variant)
        data.setup.close()
        out[key] = {"vk": vk_path, "proof": proof_path, "prove_seconds": dt}
    return out
