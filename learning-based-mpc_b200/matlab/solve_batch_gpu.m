function out = solve_batch_gpu(model, form, variant, N, dx0, dx_ref, d_off, warm)
% SOLVE_BATCH_GPU  many independent (LB)MPC QPs in one call: dx0 is nx x batch (one column per
% initial condition / scenario).  Thin wrapper over lbmpc_mex('solve').
if nargin < 6, dx_ref = []; end
if nargin < 7, d_off = []; end
if nargin < 8, warm = []; end
cfg = struct('form',form,'variant',variant,'N',N,'max_batch',size(dx0,2));
h = lbmpc_mex('create', model, cfg);
cleaner = onCleanup(@() lbmpc_mex('destroy', h));
out = lbmpc_mex('solve', h, dx0, dx_ref, d_off, warm);
end
