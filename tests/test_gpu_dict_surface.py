"""Dictionary surface on the device vs the oracle (include/obgpu_scan.h: obgpu_block_distinct_count / read_distinct /
read_reference / obgpu_filter_dict_pass / obgpu_block_group_by / obgpu_result_group_by): every dictionary-coded codec,
PAX and CS."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ora
from test_dict_surface import dict_cases, group_by_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import oceanbase_b200
    return oceanbase_b200


@pytest.fixture(scope="module")
def ctx(ob):
    c = ob.ScanContext(0)
    yield c
    c.close()


AGGS = [(0, -1), (0, 1), (1, 1), (3, 1), (4, 1), (1, 2), (4, 2)]   # COUNT(*), COUNT(c1), SUM / MIN / MAX(c1), SUM / MAX(c2)


@pytest.mark.parametrize("case", range(9))
def test_distinct_reference_black_filter_group_by(ob, ctx, case):
    name, cols = dict_cases(ob)[case]
    table = ob.encode_table(cols, 250)
    batch = ctx.open_batch(table)
    base = table.image.ctypes.data
    is_str = cols[0].obj_type == ob.OBJ_VARCHAR
    for b in range(table.n_blocks):
        blk = ora.Block(table.block(b))
        n = blk.row_count
        count = blk.dict_count(0)
        assert batch.distinct_count(b, 0) == count
        vals, lens = batch.read_distinct(b, 0, string_base=base)
        for g in range(count):
            want, off = blk.dict_entry(0, g)
            if is_str:
                assert int(vals[g]) == base + int(table.offsets[b]) + off and lens[g] == len(want)
                assert C.string_at(int(vals[g]), int(lens[g])) == want
            else:
                assert int(vals[g]) == want
        rid = np.concatenate([np.arange(0, n, 3), np.arange(n - 1, -1, -7)]).astype(np.int32)
        assert np.array_equal(batch.read_reference(b, 0, rid), blk.dict_refs(0, rid))
        # black filter: an arbitrary predicate over the distinct values ("value image is odd" / "length is even")
        verdict = np.array([(int(vals[g]) >> 9) & 1 if not is_str else (lens[g] % 2 == 0) for g in range(count)], dtype=np.uint8)
        all_refs = blk.dict_refs(0, np.arange(n, dtype=np.int32))
        for null_pass in (False, True):
            for start, cnt in ((0, None), (11, n - 30)):
                got = batch.filter_dict_pass(b, 0, verdict, null_pass, start, cnt)
                r = all_refs[start:start + (cnt if cnt is not None else n - start)]
                want = np.where(r < count, np.append(verdict, 0)[np.minimum(r, count)], int(null_pass)).astype(np.uint8)
                assert np.array_equal(got, want), (name, b, null_pass, start)
        # GROUP BY on the listed rows
        rows = np.arange(1, n, 2, dtype=np.int32)
        assert np.array_equal(batch.group_by(b, 0, AGGS, rows), group_by_model(blk, rows, AGGS)), (name, b)
    # not dictionary coded: refused like the reference's can_apply_black / group-by checks
    with pytest.raises(RuntimeError):
        batch.distinct_count(0, 1)
    with pytest.raises(RuntimeError):
        batch.filter_dict_pass(0, 0, np.zeros(batch.distinct_count(0, 0) + 1, dtype=np.uint8))
    batch.close()


@pytest.mark.parametrize("case", [0, 2, 3, 5, 7, 8])
def test_group_by_over_a_scan_result(ob, ctx, case):
    """GROUP BY pushdown for a whole page batch: the rows the white filter selected, every block in one launch."""
    name, cols = dict_cases(ob, n=3000, seed=17)[case]
    table = ob.encode_table(cols, 190)
    batch = ctx.open_batch(table)
    for flt in (None, ob.White(2, ob.WHITE_OP_LT, (1 << 30,))):
        res = batch.scan(flt, [2])
        goff, out = res.group_by(0, AGGS)
        assert goff[0] == 0 and goff[-1] == out.shape[1]
        for b in range(table.n_blocks):
            blk = ora.Block(table.block(b))
            n = blk.row_count
            rows = np.arange(n, dtype=np.int32)
            if flt is not None:
                rows = rows[blk.filter_tree(flt, 0, n).astype(bool)]
            assert goff[b + 1] - goff[b] == blk.dict_count(0) + 1
            assert np.array_equal(out[:, goff[b]:goff[b + 1]], group_by_model(blk, rows, AGGS)), (name, b)
        res.free()
    batch.close()
