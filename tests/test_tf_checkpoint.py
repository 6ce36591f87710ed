"""TensorFlow V2 checkpoint (tensor bundle) reader without TensorFlow.  No TF-written file is reachable
here, so the layers are pinned separately: CRC-32C against its published check value, varints / protobuf / table blocks
against hand-assembled byte strings that follow the LevelDB table and BundleEntryProto definitions, and the whole reader
against the writer of the same module (round trip, multi-block index, corruption detection)."""
import struct

import numpy as np
import pytest

from tacotron_b200 import tf_checkpoint as TFC


def test_crc32c_known_answers():
    assert TFC.crc32c(b"123456789") == 0xE3069283                      # the CRC-32C (Castagnoli) check value
    assert TFC.crc32c(b"") == 0
    assert TFC.crc32c(bytes(32)) == 0x8A9136AA                         # RFC 3720 B.4: 32 bytes of zeros
    assert TFC.crc32c(bytes([0xff] * 32)) == 0x62A8AB43                # RFC 3720 B.4: 32 bytes of ones
    assert TFC.mask_crc(0) == 0xa282ead8


def test_varint_and_proto_layers():
    assert TFC._put_varint(300) == b"\xac\x02" and TFC._get_varint(b"\xac\x02", 0) == (300, 2)
    # BundleEntryProto written by hand: dtype = DT_FLOAT (1), shape {dim{size:2} dim{size:3}}, offset 24, size 24, crc32c fixed32
    dim = lambda n: b"\x12" + bytes([2]) + b"\x08" + bytes([n])
    shape = dim(2) + dim(3)
    msg = b"\x08\x01" + b"\x12" + bytes([len(shape)]) + shape + b"\x20\x18" + b"\x28\x18" + b"\x35" + struct.pack("<I", 0xdeadbeef)
    m = TFC._parse_proto(msg)
    assert m[1] == [1] and m[4] == [24] and m[5] == [24] and m[6] == [0xdeadbeef]
    dims = [TFC._parse_proto(d)[1][0] for d in TFC._parse_proto(m[2][0])[2]]
    assert dims == [2, 3]
    assert TFC._entry_proto(1, (2, 3), 24, 24, 0xdeadbeef) == msg     # the writer emits exactly these bytes


def test_prefix_compressed_block_by_hand():
    # two entries "abc" -> "1", "abd" -> "22" (second shares the prefix "ab"), one restart point at 0
    block = bytes([0, 3, 1]) + b"abc" + b"1" + bytes([2, 1, 2]) + b"d" + b"22" + struct.pack("<I", 0) + struct.pack("<I", 1)
    assert list(TFC._block_entries(block)) == [(b"abc", b"1"), (b"abd", b"22")]
    assert TFC._build_block([(b"abc", b"1"), (b"abd", b"22")]) == block


@pytest.mark.parametrize("block_size", [4096, 64])                     # 64: forces many data blocks + a multi-entry index block
def test_round_trip(tmp_path, block_size):
    rng = np.random.RandomState(0)
    tensors = {"embedding/embedding": rng.randn(40, 256).astype(np.float32),
               "encoder/cbhg/conv1d/kernel": rng.randn(1, 128, 128).astype(np.float32),
               "encoder/cbhg/conv1d_1/kernel": rng.randn(2, 128, 128).astype(np.float32),
               "global_step": np.array(12345, dtype=np.int64),
               "stft_mean": rng.randn(2050).astype(np.float16),
               "stft_std": rng.rand(2050).astype(np.float32),
               "scalar_f": np.array(2.5, dtype=np.float32)}
    for i in range(30):
        tensors[f"decoder/decoder/var_{i:02d}/Adam_1"] = rng.randn(3, i + 1).astype(np.float32)
    prefix = str(tmp_path / "tacotron-12345")
    TFC.write_bundle(prefix, tensors, block_size=block_size)
    header, entries = TFC.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and set(entries) == set(tensors)
    got = TFC.read_bundle(prefix)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    assert TFC.latest_checkpoint(str(tmp_path)) == prefix
    sub = TFC.read_bundle(prefix, names={"global_step"})
    assert list(sub) == ["global_step"] and int(sub["global_step"]) == 12345


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "m-1")
    TFC.write_bundle(prefix, {"a": np.arange(10, dtype=np.float32)})
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[5] ^= 0x10
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    with pytest.raises(TFC.BundleError):
        TFC.read_bundle(prefix)
    assert np.array_equal(TFC.read_bundle(prefix, verify_crc=False)["a"][2:], np.arange(10, dtype=np.float32)[2:])
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[-1] ^= 0xff                                                     # footer magic
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(TFC.BundleError):
        TFC.read_index(prefix + ".index")


def test_import_maps_tf_names(tmp_path):
    """a bundle with the reference's TF variable names loads through checkpoint.import_tf_checkpoint"""
    import torch
    from tacotron_b200 import checkpoint
    loaded = {}

    class FakeModel:
        global_step = 0
        def load_params(self, p):
            loaded.update(p)
    w = np.random.RandomState(1).randn(40, 256).astype(np.float32)
    prefix = str(tmp_path / "tacotron-7")
    TFC.write_bundle(prefix, {"embedding/embedding": w, "global_step": np.array(7, dtype=np.int64),
                              "stft_mean": np.zeros(4, np.float16)})
    m = FakeModel()
    extra = checkpoint.import_tf_checkpoint(m, prefix)
    assert torch.equal(loaded["embedding"], torch.from_numpy(w)) and m.global_step == 7 and "stft_mean" in extra
