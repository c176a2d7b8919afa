function reference_dump(refdir, infile, outfile)
% REFERENCE_DUMP  run the UNMODIFIED reference solver under MATLAB / Octave with injected initial factors and save
% its outputs for tests/test_reference_pin.py::test_oracle_matches_matlab_dump (SURVEY.md 8c).
%   python tools/reference_dump_inputs.py                      % writes tests/golden/dump_in_<case>.mat
%   octave --eval "addpath('tools'); reference_dump('/path/to/reference', 'tests/golden/dump_in_cfg1.mat', 'tests/golden/matlab_cfg1.mat')"
% randn (fast_robust_triple_tensor/triple_decomp_ADMM.m:23) is shadowed by tools/shadow/randn.m, which hands out
% A0, B0, C0 in call order; nothing else of the reference is touched.
    S = load(infile);                                   % D, r, opts, A0, B0, C0
    global TRITD_RANDN_QUEUE
    TRITD_RANDN_QUEUE = {S.A0, S.B0, S.C0};
    addpath(fullfile(refdir, 'fast_robust_triple_tensor'));
    addpath(fullfile(fileparts(mfilename('fullpath')), 'shadow'));
    D = S.D; r = double(S.r); opts = S.opts; A0 = S.A0; B0 = S.B0; C0 = S.C0;
    t = tic;
    [A, B, C, O, errHist] = triple_decomp_ADMM(D, r, opts);
    seconds = toc(t); iters = numel(errHist);
    fprintf('reference_dump: %d iterations in %.3f s (%.3f it/s)\n', iters, seconds, iters / seconds);
    save(outfile, 'D', 'r', 'opts', 'A0', 'B0', 'C0', 'A', 'B', 'C', 'O', 'errHist', 'seconds', '-v7');
end
