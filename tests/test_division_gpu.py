"""GPU: the shared-reciprocal division of frag_weights_fast() (swgl_dev_math.cuh) against the
correctly rounded `/` the reference performs (swgl.c:3256-3268, 3367-3376), on the operand domains
the fast path is restricted to.  Bit-exact: zero mismatches allowed."""
import pytest

import swgl_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pairs", [1 << 20, (1 << 28) + 12345])
def test_shared_reciprocal_division_is_correctly_rounded(pairs):
    api = swgl_b200.load()
    api.glInit(64, 64)
    assert not api.swglGetLastError()
    api.swglSetOption(b"selftest_division", pairs)
    assert api.swglGetOption(b"selftest_division_mismatches") == 0
