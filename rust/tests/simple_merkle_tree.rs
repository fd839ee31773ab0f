//! The reference's own tests for the simple tree (/root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:117-310),
//! restated against the GPU wrapper `build_gpu`: same leaves, same known answers (:136-140, :181-190, :210-211), the
//! reference's own `get_merkle_proof` / `verify_merkle_proof` running unchanged on the tree the engine built, and the
//! engine's tree compared with the reference's CPU `MerkleTree::build` field by field.
//! UNCOMPILED in this repository's image; run with `cargo test -- --test-threads=1` on a B200 box.
use plonky2::field::goldilocks_field::GoldilocksField as F;
use plonky2::field::types::Field;
use plonky2::hash::hash_types::HashOut;
use plonky2_merkle_trees::simple_merkle_tree::simple_merkle_tree::{verify_merkle_proof, MerkleTree};
use pmt_shim::pmt_ffi::Ctx;
use pmt_shim::wrappers::build_gpu;

fn h(e: [u64; 4]) -> HashOut<F> {
    HashOut { elements: [F::from_canonical_u64(e[0]), F::from_canonical_u64(e[1]), F::from_canonical_u64(e[2]), F::from_canonical_u64(e[3])] }
}
fn leaves4() -> Vec<F> {
    [2890852870u64, 156728478, 2876514289, 984286162].iter().map(|&x| F::from_canonical_u64(x)).collect()
}
fn leaves16() -> Vec<F> {
    [14786323743454721611u128, 976503040092093812, 4644130751253292674, 6522877527545910706, 11021172818651636092,
     12048403458499719587, 11457874926809001558, 14982007443548219923, 4546369223935415035, 7205140577604465038,
     4644130751253292674, 4208177174652750506, 16147116534354400672, 18147003476480002882, 14133393155459789216,
     9890944065319669426].iter().map(|&x| F::from_noncanonical_u128(x)).collect()
}
fn same(a: &MerkleTree, b: &MerkleTree) {
    assert_eq!(a.count_levels, b.count_levels);
    assert!(a.tree == b.tree && a.root == b.root);
}

#[test]
fn test_build_merkle_tree_4_leaves() {
    let ctx = Ctx::new(0);
    let tree = build_gpu(&ctx, leaves4());
    assert_eq!(tree.count_levels, 2);
    assert!(tree.tree[1][0] == h([6678006133445961348, 15827935749738443865, 6295652393730592048, 1546515167911236130]));   // :138
    assert!(tree.tree[1][1] == h([6698018865469624861, 12486244005715193285, 11330639022572315007, 6059804404595156248]));
    assert!(tree.root == h([13451271846715771774, 4069913004933160254, 14528216580130305557, 9716424959297545638]));         // :140
    same(&tree, &MerkleTree::build(leaves4()));
}

#[test]
fn test_build_merkle_tree_16_leaves() {
    let ctx = Ctx::new(0);
    let tree = build_gpu(&ctx, leaves16());
    assert_eq!(tree.count_levels, 4);
    assert!(tree.tree[1][0] == h([16072672881132969138, 16679487992876356669, 4319836168073005766, 14599992432910949662]));  // :184
    assert!(tree.tree[2][3] == h([9702041242754623164, 9442892912940285811, 2205638039663440432, 4535189628500499303]));      // :186
    assert!(tree.tree[3][1] == h([14079844864384152521, 6499705357519308869, 16026207645313349904, 15079809878245341298]));   // :188
    assert!(tree.root == h([2659148958598424285, 16496267010313658247, 12216516055477211974, 15749220035779350537]));         // :190
    same(&tree, &MerkleTree::build(leaves16()));
}

#[test]
fn test_merkle_proof_small_tree() {
    let ctx = Ctx::new(0);
    let tree = build_gpu(&ctx, leaves4());
    let res_leaf_0 = tree.clone().get_merkle_proof(0);                                   // the reference's own method
    assert!(res_leaf_0[0] == h([156728478, 0, 0, 0]));                                    // :210
    assert!(res_leaf_0[1] == h([6698018865469624861, 12486244005715193285, 11330639022572315007, 6059804404595156248]));   // :211
}

#[test]
fn test_verify_small_merkle_proof() {
    let ctx = Ctx::new(0);
    let leaves = leaves4();
    let tree = build_gpu(&ctx, leaves.clone());
    assert!(verify_merkle_proof(leaves[0], 0, tree.root, tree.clone().get_merkle_proof(0)));
    assert!(verify_merkle_proof(leaves[3], 3, tree.root, tree.clone().get_merkle_proof(3)));
}

#[test]
fn test_verify_merkle_proof_16() {
    let ctx = Ctx::new(0);
    let leaves = leaves16();
    let tree = build_gpu(&ctx, leaves.clone());
    let proofs: Vec<_> = (0..16).map(|i| tree.clone().get_merkle_proof(i)).collect();
    for i in 0..16 {
        assert!(verify_merkle_proof(leaves[i], i, tree.root, proofs[i].clone()));
    }
    assert!(!verify_merkle_proof(leaves[1], 0, tree.root, proofs[0].clone()));            // wrong leaf  (:298)
    assert!(!verify_merkle_proof(leaves[0], 1, tree.root, proofs[0].clone()));            // wrong index (:300)
    assert!(!verify_merkle_proof(leaves[0], 0, tree.root, proofs[1].clone()));            // wrong proof (:303)
    assert!(!verify_merkle_proof(leaves[0], 0, tree.tree[0][0], proofs[0].clone()));      // wrong root  (:306)
}

#[test]
#[should_panic]
fn build_rejects_a_non_power_of_two_like_log2_strict() {
    let ctx = Ctx::new(0);
    let _ = build_gpu(&ctx, leaves4()[..3].to_vec());                                     // simple_merkle_tree.rs:30
}
