/* mbd_pusht.h — parameter and state layout of the pushT env (the one env of the reference on Brax's `generalized`
 * backend: /root/reference/mbd/envs/pushT.py:16-20, model /root/reference/mbd/assets/pushT.xml).
 *
 * Shared by the CUDA kernel (mbd_b200/csrc/pusht.cuh), the CPU oracle (oracle/pusht_oracle.c) and the host env
 * (mbd_b200/envs/pusht.py): constants only, no code.
 *
 * The model is three world-parented planar bodies in reduced coordinates, q = [pusher x, y | slider x, y, theta |
 * goal x, y, theta], no gravity:
 *   pusher  sphere, two actuated slide dofs (motor gear 30, ctrl clipped to [-1, 1]);
 *   slider  two boxes welded into a T, slide x, slide y, hinge z with joint damping; collides with the pusher;
 *   goal    a ghost T (no collisions, no forces): its coordinates only enter the reward.
 * One physics step is Brax's generalized pipeline [brax-recalled, UNPINNED — see DESIGN.md]: smooth forces (actuation, joint
 * damping, centrifugal bias), soft constraints in MuJoCo's solref / solimp parameterisation (joint limits, sphere-box contacts
 * with a 4-sided friction pyramid), a projected Gauss-Seidel solve of the constraint QP, semi-implicit Euler with the joint
 * damping folded into the mass matrix.
 */
#ifndef MBD_PUSHT_H_
#define MBD_PUSHT_H_

#define MBD_PT_NQ 8        /* generalized coordinates */
#define MBD_PT_STATE 16    /* q[8] | qd[8] */
#define MBD_PT_NU 2
#define MBD_PT_NBOX 2
#define MBD_PT_NLIM 4      /* limited slide dofs: q0 q1 (pusher), q2 q3 (slider) */
#define MBD_PT_NCROW 3     /* rows per contact: n - mu t, n + mu t, and the out-of-plane pyramid pair merged into one n row */
#define MBD_PT_NROW (MBD_PT_NLIM + MBD_PT_NCROW * MBD_PT_NBOX)

enum {
  MBD_PT_DT = 0,    /* physics step (option timestep) */
  MBD_PT_NSUB,      /* n_frames of the env (pushT.py:20), stored as float */
  MBD_PT_ITERS,     /* constraint-solver sweeps (MuJoCo opt.iterations default 100), stored as float */
  MBD_PT_GEAR0, MBD_PT_GEAR1,
  MBD_PT_MP, MBD_PT_IMP, MBD_PT_RP, MBD_PT_DPX, MBD_PT_DPY,      /* pusher: mass, 1/mass, radius, joint damping */
  MBD_PT_MS, MBD_PT_IMS, MBD_PT_IS, MBD_PT_IIS,                  /* slider: mass, 1/mass, Izz about the COM, 1/Izz */
  MBD_PT_CX, MBD_PT_CY,                                          /* slider COM in the body frame */
  MBD_PT_DSX, MBD_PT_DSY, MBD_PT_DSTH,                           /* slider joint damping */
  MBD_PT_LIM0,                                                   /* (lo, hi) x MBD_PT_NLIM */
  MBD_PT_BOX0 = MBD_PT_LIM0 + 2 * MBD_PT_NLIM,                   /* (centre x, y, half x, half y) x MBD_PT_NBOX, body frame */
  MBD_PT_MU = MBD_PT_BOX0 + 4 * MBD_PT_NBOX,                     /* friction coefficient of the pair */
  MBD_PT_DMIN, MBD_PT_DMAX, MBD_PT_WIDTH, MBD_PT_MID,            /* solimp (power must be 2) */
  MBD_PT_KB, MBD_PT_KK,                                          /* solref: b = 2/(dmax tc), k = 1/(dmax^2 tc^2 dr^2) */
  MBD_PT_TOL,       /* solver: stop after a sweep that moves the constraint force J^T x by <= TOL * its largest component */
  MBD_PT_NPARAM
};

#endif /* MBD_PUSHT_H_ */
