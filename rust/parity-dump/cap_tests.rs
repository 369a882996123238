// Paste into the `mod test` of a CAP checkout (jf-cap 0.0.4):
//   * this block into src/proof/transfer.rs   (next to test_transfer_validity_proof, :600),
//   * the analogous two-liner into src/proof/mint.rs (:345) and src/proof/freeze.rs (:430),
// add `capgpu-parity-dump = { path = ".../capgpu/rust/parity-dump" }` under [dev-dependencies], run
//   cargo test --release --features "bn254" dump_capgpu_fixture -- --nocapture
// and copy target/upstream_*.capfix into capgpu/tests/fixtures/.  The test repeats the body of
// `super::prove` (src/proof/transfer.rs:159-188) so that the circuit and the key reach the dumper; it
// then verifies the proof exactly like test_transfer_validity_proof does.
#[test]
fn dump_capgpu_fixture() -> Result<(), TxnApiError> {
    use ark_serialize::CanonicalSerialize;
    let rng = &mut ark_std::test_rng();
    let (num_input, num_output, depth) = (2usize, 2usize, 10u8); // BASELINE config 1: domain 2^15
    let max_degree = 32770;
    let universal_param = universal_setup_for_staging::<_, Config>(max_degree, rng)?;
    let (proving_key, verifying_key, _) = super::preprocess::<Config>(&universal_param, num_input, num_output, depth)?;
    let recv_memos_ver_key = schnorr::KeyPair::generate(rng).ver_key();
    let extra_proof_bound_data = "some random data".as_bytes();
    let user_keypair1 = UserKeyPair::generate(rng);
    let user_keypair2 = UserKeyPair::generate(rng);
    let builder = TransferParamsBuilder::new_non_native(num_input, num_output, Some(depth), vec![&user_keypair1, &user_keypair2])
        .set_input_amounts(30u64.into(), &Amount::from_vec(&[25])[..])
        .set_output_amounts(19u64.into(), &Amount::from_vec(&[36])[..])
        .set_input_creds(9998u64);
    let witness = builder.build_witness(rng);
    let pub_input = TransferPublicInput::from_witness(&witness, 1234u64)?;
    // src/proof/transfer.rs:167-181, with the prove call routed through the dumper
    let (circuit, _) = TransferCircuit::build(&witness, &pub_input).map_err(|e| TxnApiError::FailedSnark(format!("{:?}", e)))?;
    circuit.0.check_circuit_satisfiability(&pub_input.to_scalars()).map_err(|e| TxnApiError::FailedSnark(format!("{:?}", e)))?;
    let mut ext_msg = Vec::new();
    CanonicalSerialize::serialize(&recv_memos_ver_key, &mut ext_msg)?;
    ext_msg.extend_from_slice(extra_proof_bound_data);
    let proof = capgpu_parity_dump::dump_fixture(
        std::path::Path::new("target/upstream_transfer_2x2_depth10.capfix"),
        [0, num_input as u64, num_output as u64, depth as u64],
        ark_std::test_rng(), // a FRESH test_rng: the replay starts from the same stream position
        &circuit.0, &proving_key.proving_key, Some(ext_msg),
    ).map_err(|e| TxnApiError::FailedSnark(format!("{:?}", e)))?;
    assert!(super::verify(&verifying_key, &pub_input, &proof, &recv_memos_ver_key, extra_proof_bound_data).is_ok());
    Ok(())
}
