function solver = solver_gpu(model, N, variant, x_eq, u_eq, delta)
% SOLVER_GPU  stand-in for the CasADi object built at DMS_tracking_LMPC_casadi.m:126-127
%   solver = nlpsol('solver','ipopt',nlp)
% with the same name/value call convention as DMS_tracking_LMPC_casadi.m:163-167:
%   res = solver('x0',y_init,'lbx',lb,'ubx',ub,'lbg',con_lb,'ubg',con_ub);
%   y_OL = full(res.x);   res.f
% y = [x_0..x_N ; u_0..u_{N-1} ; theta] in ABSOLUTE coordinates like the reference's `y`
% (DMS_tracking_LMPC_casadi.m:122,168-172).  The measured state is read from lbx(1:n) (= ubx(1:n),
% :161-162); the guess x0 seeds the inputs/theta; lbg/ubg are implied by `model` and ignored.
if nargin < 6, delta = 0.01; end
cfg = struct('form','C','variant',variant,'N',N,'delta',delta,'max_batch',1);
h = lbmpc_mex('create', model, cfg);
n = size(model.A,1); m = size(model.B,2); nt = size(model.LAMBDA,2);
keep = onCleanup(@() lbmpc_mex('destroy', h)); %#ok<NASGU>
solver = @call;
    function res = call(varargin)
        keep; %#ok<VUNUS> ties the handle's lifetime to this closure
        a = struct(varargin{:});
        dx0 = a.lbx(1:n) - x_eq;
        warm = [];
        if isfield(a,'x0') && ~isempty(a.x0)
            y0 = a.x0(:);
            du0 = y0(n*(N+1)+1:end-nt) - repmat(u_eq, N, 1);
            warm = [du0; y0(end-nt+1:end)];
        end
        out = lbmpc_mex('solve', h, dx0, [], [], warm);
        xabs = reshape(out.x, n, N+1) + repmat(x_eq, 1, N+1);
        res.x = [xabs(:); out.u_or_c + repmat(u_eq, N, 1); out.theta];
        res.f = out.f;
        res.status = out.status; res.iters = out.iters;
    end
end

function y = full(x) %#ok<DEFNU> (CasADi's full(); doubles pass through)
y = x;
end
