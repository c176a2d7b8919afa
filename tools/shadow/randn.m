function X = randn(varargin)
% shadow of randn for tools/reference_dump.m: returns the queued initial factors in call order (A0, B0, C0) and
% checks the sizes the reference asks for (triple_decomp_ADMM.m:23: randn(n1,r,r), randn(r,n2,r), randn(r,r,n3))
    global TRITD_RANDN_QUEUE
    X = TRITD_RANDN_QUEUE{1};
    TRITD_RANDN_QUEUE(1) = [];
    want = [varargin{:}];
    sz = size(X); sz(end + 1:numel(want)) = 1;
    assert(isequal(sz, want), 'randn shadow: size mismatch');
end
