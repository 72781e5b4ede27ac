function [sysHistory,art_refHistory,true_refHistory]...
          =ocpLBMPC_gpu(x,x_wp,dx_init,dx_ref,u_wp,...
                    N,Ts,iterations,options,opt_var,data,...
                    A,B,Kstabil,Q,R,P,T,Mtheta,LAMBDA,PSI,m,...
                    F_x,h_x,F_u,h_u,F_w_N,h_w_N,F_x_d,h_x_d,...
                    sysHistory,art_refHistory,true_refHistory)
% OCPLBMPC_GPU  drop-in for functions/ocpLBMPC.m (same 35 arguments, same 3 outputs).
% The only change is the solver call: the fmincon-SQP solve of ocpLBMPC.m:27-31 becomes one
% batched interior-point solve on the GPU (batch = 1 here) through lbmpc_mex -> liblbmpc_b200.so.
% `options` (fmincon options) is accepted and ignored.
%
% The reference's problem is stated as the reference states it: costLBMPC.m:27 rolls the LEARNED model
% (transitionLearned.m:13-14: u = K x + c on the learned state, x+ = A x + B u + g(x,u;data)) while
% constraintsLBMPC.m:23 rolls the NOMINAL model, so the oracle moves the cost only and the X, U, X(-)D
% and terminal rows stay on the nominal prediction.  The non-convex oracle term is linearised along the
% previous solution `sqp_iters` times per control step (lbmpc_solve_sqp_ex with twin = 1, order = 1: oracle
% value AND Jacobian give an LTV QP on the learned sequence whose rows follow the nominal one at the gap
% e+ = (A + B K) e + g_k, input gap K e) — the same call the Python mirror lbmpc_b200.ocpLBMPC makes, which
% reproduces the reference's saved LBMPC_N50_sys_full.mat history to 5e-6 on the inputs (tests/).
% With data = 0 (first step) this is the exact QP of the reference.
model = struct('A',A,'B',B,'K',Kstabil,'Q',Q,'R',R,'P',P,'T',T,'LAMBDA',LAMBDA,'PSI',PSI, ...
               'F_x',F_x,'h_x',h_x,'F_u',F_u,'h_u',h_u,'F_w_N',F_w_N,'h_w_N',h_w_N, ...
               'F_x_d',F_x_d,'h_x_d',h_x_d);
cfg = struct('form','F','variant','LBMPC','N',N,'max_batch',1);
h = lbmpc_mex('create', model, cfg);
cleaner = onCleanup(@() lbmpc_mex('destroy', h));
q = 100;                                   % moving window (ocpLBMPC.m:18)
sqp_iters = 2;                             % linearisations of the oracle per control step (drivers.py default)
for k = 1:iterations
    if k > 1
        X = [x(1:2)-x_wp(1:2); u-u_wp];                        % ocpLBMPC.m:14
        Y = (x_k1-x_wp) - (A*(x-x_wp) + B*(u-u_wp));           % ocpLBMPC.m:15
        x = x_k1;
        data = update_data(X,Y,q,k,data);                      % ocpLBMPC.m:19
        dx = x - x_wp;
    else
        dx = dx_init;
    end
    % replaces fmincon(COSTFUN,opt_var,...,CONSFUN,options), ocpLBMPC.m:27-31
    out = lbmpc_mex('solve_sqp', h, sqp_iters, size(data.X,2), 0.5, 0.001, dx, dx_ref, data.X, data.Y, [], opt_var(:), 1, 1);   % twin = 1, order = 1
    if out.status ~= 0
        warning('lbmpc:status', 'step %d: solver status %d', k, out.status);
    end
    opt_var = [out.u_or_c; out.theta];
    theta_opt = reshape(opt_var(end-m+1:end), m, 1);
    c = reshape(opt_var(1:m), m, 1);
    art_ref = Mtheta*theta_opt;
    [x_k1, u] = transitionTrue(x, c, x_wp, u_wp, Kstabil, Ts);   % plant stays in MATLAB (ocpLBMPC.m:37)
    his = [x-x_wp; u-u_wp];
    sysHistory = [sysHistory his]; %#ok<*AGROW>
    art_refHistory = [art_refHistory art_ref(1:m)];
    true_refHistory = [true_refHistory dx_ref];
end
end
