function out = solve_learned_gpu(model, N, dx0, X, Y, varargin)
% SOLVE_LEARNED_GPU  learned-oracle LBMPC (costLBMPC.m:27, DMS_LBMPC_casadi.m:252-319) as a sequence of QPs.
%   out = solve_learned_gpu(model, N, dx0, X, Y, 'twin', true, 'sqp_iters', 3, 'bandwidth', 0.5, 'lambda', 1e-3, ...
%                           'dx_ref', [], 'valid', [], 'warm', [])
%   dx0 nx x batch; X 3 x q x batch, Y nx x q x batch: the data windows (data.X, data.Y of update_data.m, one per QP);
%   valid q x batch: the mask of casadiL2NW.m:18-21.
%   twin = false: one state sequence, the oracle correction enters the dynamics (LBMPC_casadi.m);
%   twin = true : the reference's two sequences — cost on the learned states, rows on the nominal ones.
%   out: u_or_c, theta, x (nominal states when twin), f, iters, status, du_step (sqp_iters x batch).
p = inputParser;
addParameter(p, 'twin', false); addParameter(p, 'sqp_iters', 3); addParameter(p, 'bandwidth', 0.5);
addParameter(p, 'lambda', 1e-3); addParameter(p, 'dx_ref', []); addParameter(p, 'valid', []); addParameter(p, 'warm', []);
parse(p, varargin{:});
o = p.Results;
cfg = struct('form','C','variant','LBMPC','N',N,'max_batch',size(dx0,2));
h = lbmpc_mex('create', model, cfg);
cleaner = onCleanup(@() lbmpc_mex('destroy', h));
out = lbmpc_mex('solve_sqp', h, o.sqp_iters, size(X,2), o.bandwidth, o.lambda, dx0, o.dx_ref, X, Y, o.valid, o.warm, double(o.twin));
end
