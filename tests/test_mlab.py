"""Unit tests of the MATLAB-subset interpreter (oracle/mlab.py, test infrastructure): the MATLAB semantics the
reference's sources rely on -- column-major reshape / permute, leading-dimension broadcasting, indexing and indexed
assignment, whitespace rules in matrix literals, multiple outputs, local functions shadowing the path."""
import numpy as np
import pytest

import mlab


def run(src, name, args, nargout=1, tmp_path=None, files=None):
    d = tmp_path
    (d / (name + ".m")).write_text(src)
    for fn, text in (files or {}).items():
        (d / fn).write_text(text)
    outs, text, it = mlab.run_function([str(d)], name, args, nargout=nargout)
    return outs, text, it


def test_reshape_permute_are_column_major(tmp_path):
    src = "function [a, b, c] = f(X)\n a = reshape(X, [3, 8]);\n b = permute(X, [3 1 2]);\n c = reshape(permute(X, [2, 1, 3]), [3, 2*4]);\nend\n"
    X = np.arange(24.0).reshape((2, 3, 4), order="F")
    (a, b, c), _, _ = run(src, "f", [X], 3, tmp_path)
    assert np.array_equal(a, X.reshape((3, 8), order="F"))
    assert np.array_equal(b, np.transpose(X, (2, 0, 1)))
    assert np.array_equal(c, np.transpose(X, (1, 0, 2)).reshape((3, 8), order="F"))


def test_broadcast_aligns_leading_dimensions(tmp_path):
    src = "function F = f(B, C)\n F = reshape(B, [4, 6, 1]) .* reshape(C', [1, 6, 5]);\nend\n"
    B = np.random.default_rng(0).standard_normal((4, 6)); C = np.random.default_rng(1).standard_normal((5, 6))
    (F,), _, _ = run(src, "f", [B, C], 1, tmp_path)
    assert F.shape == (4, 6, 5)
    assert np.allclose(F, B[:, :, None] * C.T[None, :, :])


def test_indexing_and_assignment(tmp_path):
    src = ("function [a, b, c, d, e, v] = f(X, h)\n a = X(:);\n b = X(2,:);\n c = h(2:end);\n"
           " A = zeros(3, 2, 2);\n for i = 1:3\n  A(i,:,:) = reshape(X(i,:), [2, 2]);\n end\n d = A;\n"
           " e = X; e(2, 3) = -1; e(end) = 7;\n v = h; v(2) = 5;\nend\n")
    X = np.arange(12.0).reshape((3, 4), order="F"); h = np.array([[1.0], [2.0], [3.0]])
    (a, b, c, d, e, v), _, _ = run(src, "f", [X, h], 6, tmp_path)
    assert a.shape == (12, 1) and np.array_equal(a.ravel(), X.ravel(order="F"))
    assert b.shape == (1, 4) and np.array_equal(b.ravel(), X[1])
    assert c.shape == (2, 1) and np.array_equal(c.ravel(), [2, 3])          # vector(range) keeps the vector's orientation
    assert np.array_equal(d, X.reshape((3, 2, 2), order="F"))
    assert e[1, 2] == -1 and e[2, 3] == 7
    assert X[1, 2] == 7.0 and X[2, 3] == 11.0                              # the input is untouched: value semantics
    assert np.array_equal(v.ravel(), [1, 5, 3])


def test_matrix_literal_whitespace_and_transpose_vs_string(tmp_path):
    src = ("function [a, b, c, s] = f(x)\n a = [x -1];\n b = [x - 1];\n c = [x' x'];\n s = sprintf('%d-%s', 3, 'it''s');\nend\n")
    (a, b, c, s), _, _ = run(src, "f", [np.array([[2.0, 4.0]])], 4, tmp_path)
    assert np.array_equal(a, [[2, 4, -1]]) and np.array_equal(b, [[1, 3]])
    assert c.shape == (2, 2) and np.array_equal(c, [[2, 2], [4, 4]])
    assert s == "3-it's"


def test_size_outputs_tilde_and_struct_fields(tmp_path):
    src = ("function [n1, n3, m, q] = f(X, opts)\n [n1, ~, n3] = size(X);\n [~, m] = size(X);\n q = opts.mu*1e6;\n"
           " if opts.disp && mod(10,10)==0\n  fprintf(\"Iter %d, errL=%.2e\\n\", 10, 0.00123);\n end\nend\n")
    outs, text, _ = run(src, "f", [np.zeros((2, 3, 4)), dict(mu=1e-3, disp=1)], 4, tmp_path)
    assert [mlab.scalar(x) for x in outs] == [2, 4, 12, 1e-3 * 1e6]
    assert text == "Iter 10, errL=1.23e-03\n"
    with pytest.raises(mlab.MlabError, match='Unrecognized field name "rho"'):
        run("function q = g(o)\n q = o.rho;\nend\n", "g", [dict(mu=1.0)], 1, tmp_path)


def test_local_functions_shadow_the_path_and_break(tmp_path):
    main = ("function [y, k] = f(x)\n y = helper(x);\n for k = 1:10\n  if k > 1 && abs(k - 3) < 0.5\n   break;\n  end\n end\nend\n"
            "function y = helper(x)\n y = x + 1;\nend\n")
    (y, k), _, it = run(main, "f", [1.0], 2, tmp_path, files={"helper.m": "function y = helper(x)\n y = x + 100;\nend\n"})
    assert mlab.scalar(y) == 2 and mlab.scalar(k) == 3
    assert [n for n, _ in it.calls] == ["f", "helper"] and it.calls[1][1].endswith("f.m")


def test_pinv_cutoff_and_switch_error(tmp_path):
    src = "function P = f(G)\n P = pinv(G);\nend\n"
    (P,), _, _ = run(src, "f", [np.diag([1.0, 1e-20, 2.0])], 1, tmp_path)
    assert P[1, 1] == 0.0 and P[0, 0] == 1.0 and P[2, 2] == 0.5
    src = "function y = g(m)\n switch m\n  case 1\n   y = 10;\n  otherwise\n   error('Mode must be 1, 2, or 3.');\n end\nend\n"
    with pytest.raises(mlab.MlabError, match="Mode must be 1, 2, or 3."):
        run(src, "g", [4.0], 1, tmp_path)
    with pytest.raises(mlab.MlabError):
        run("function y = h(x)\n y = x(0);\nend\n", "h", [np.ones((2, 2))], 1, tmp_path)
