"""Host-side pieces of bench.py that need no GPU: the CSV export in the column layout of the reference's benches
(/root/reference/src/bench_utils/mod.rs:236-253, written by save_result_to_file_simple)."""
import csv
import types

import bench


def test_reference_csv_layout(tmp_path):
    line = {"value": 1000.0, "n_gpus": 2, "config": {"host_threads_per_gpu": 4},
            "configs": {"note_shapes": {"mint": {"domain": "2^14", "ms_per_proof": 1.1}, "freeze_5": {"domain": "2^16", "ms_per_proof": 3.9},
                                        "transfer_2x2_batch_1024": {"domain": "2^15", "ms_per_proof": 2.0}}}}
    path = tmp_path / "transfer_note_cap_benchmark.csv"
    bench.write_reference_csv(str(path), line, types.SimpleNamespace(log_n=15), types.SimpleNamespace(workload="transfer_2x2", ctxs=4))
    rows = list(csv.reader(open(path)))
    assert rows[0] == ["TRANSACTION", "N_THREADS", "FUNCTION", "N_INPUTS", "N_OUTPUTS", "TREE_HEIGHT", "DOMAIN_SIZE", "N_CONSTRAINTS",
                       "UTILITY_RATIO(%)", "TRANSFER_NOTE_SIZE (KB)", "PROVING_KEY_SIZE (KB)", "VERIFYING_KEY_SIZE (KB)", "TIME (ms)", "N_GPUS", "ROOFLINE_FRAC"]
    assert [r[0] for r in rows[1:]] == ["transfer_note", "mint_note", "freeze_note"]
    assert rows[1][2] == "Gen" and rows[1][6] == "32768" and float(rows[1][12]) == 2.0 and rows[1][13] == "2"  # 1000 proofs/s on 2 GPUs = 2 ms per proof per GPU
    assert rows[2][6] == "16384" and float(rows[2][12]) == 1.1
    assert all(len(r) == 15 for r in rows)
