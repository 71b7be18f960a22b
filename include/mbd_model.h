/* mbd_model.h — layout of the compiled model blob ("System" image) handed across the
 * C ABI (mbd_model_create) and staged into shared memory by the rollout kernel.
 *
 * It carries what Brax's `System` carries for the positional pipeline
 * (brax.io.mjcf.load at /root/reference/mbd/envs/humanoidrun.py:15): link tree, link
 * transforms, joint frames, masses, COMs, dof parameters, actuator map, the static
 * sphere-plane contact list and the solver scalars from <custom><numeric>.
 *
 * Layout: a header of MBD_HDR_WORDS 32-bit words followed by a field-major (SoA) table
 * link_field[f][l], f < MBD_NFIELDS, l < MBD_MAXL.  SoA because lane l of a sample group
 * reads field f of link l: consecutive lanes hit consecutive banks, the samples sharing a
 * warp broadcast.  Every entry is 4 bytes (float or int32, see the enum).
 * Constants are derived in float64 by mbd_b200/model/blob.py and rounded once.
 */
#ifndef MBD_MODEL_H_
#define MBD_MODEL_H_

#include <stdint.h>

#define MBD_MODEL_MAGIC 0x4D424431 /* "MBD1" */
#define MBD_MAXL 16                /* max links (lanes per sample group) */
#define MBD_MAXCHILD 4
#define MBD_MAXDOF 3
#define MBD_MAXCON 6               /* sphere-plane contacts per link (capsules contribute their two end caps) */
#define MBD_MAXTRACK 8
#define MBD_DOF_STRIDE 8
#define MBD_CON_STRIDE 5

/* ---- header words --------------------------------------------------------------- */
enum {
  MBD_H_MAGIC = 0,
  MBD_H_NLINK,        /* int  */
  MBD_H_NU,           /* int  action size */
  MBD_H_NFRAMES,      /* int  physics substeps per env step (PipelineEnv n_frames) */
  MBD_H_REWARD,       /* int  MBD_REWARD_* */
  MBD_H_NTRACK,       /* int  number of tracked bodies for eval_xref_logpd (0 = none) */
  MBD_H_TRACK0,       /* int[MBD_MAXTRACK] link ids */
  MBD_H_DT = MBD_H_TRACK0 + MBD_MAXTRACK, /* float sys.dt (opt.timestep) */
  MBD_H_INV_DT,       /* float */
  MBD_H_GX, MBD_H_GY, MBD_H_GZ,
  MBD_H_VEL_DAMP,     /* float exp(vel_damping*dt) */
  MBD_H_ANG_DAMP,     /* float exp(ang_damping*dt) */
  MBD_H_SCALE_POS,    /* float joint_scale_pos */
  MBD_H_SCALE_ANG,    /* float joint_scale_ang */
  MBD_H_COLLIDE_SCALE,
  MBD_H_ELASTICITY,
  MBD_H_HALF_DT,      /* float 0.5*dt */
  MBD_H_TWO_INV_DT,   /* float 2/dt */
  MBD_H_RW0,          /* float[4] reward parameters */
  MBD_HDR_WORDS = 64
};

enum { MBD_REWARD_HUMANOIDRUN = 0, MBD_REWARD_HUMANOIDTRACK = 1,
       MBD_REWARD_HOPPER = 2,  /* x - 0.5 clip(|z - RW0|, -1, 1): hopper (RW0 = 1.0) and walker2d (RW0 = 1.1) */
       MBD_REWARD_HUMANOIDSTANDUP = 3,
       MBD_REWARD_ANT = 4,     /* (x' - x)/RW0 + RW1 - RW2 |a|^2: ant (RW1 healthy 1, RW2 0.5) and halfcheetah (RW1 0, RW2 0.1); RW0 = env dt */
       MBD_REWARD_CARTPOLE = 5 /* cos(q[1]) - |qd[0]|: pole angle about its hinge axis, cart slide velocity (cartpole.py:44) */ };

/* ---- per-link fields ------------------------------------------------------------- */
enum {
  MBD_F_PARENT = 0,    /* int  (-1 = world) */
  MBD_F_NDOF,          /* int  0 = free root, 1..3 stacked 1-dof joints (hinges; slides only on world-parented links) */
  MBD_F_CHILD0,        /* int[MBD_MAXCHILD] ascending link ids, -1 = none */
  MBD_F_MASS = MBD_F_CHILD0 + MBD_MAXCHILD,
  MBD_F_INV_MASS,
  MBD_F_PINV_MASS,     /* inverse mass of the parent (0 for world) */
  MBD_F_PINV_INERTIA,  /* inverse (isotropic) inertia of the parent: 1, or 0 for world */
  MBD_F_COM,           /* float[3] inertia.transform.pos (link frame) */
  MBD_F_RC = MBD_F_COM + 3,   /* float[3] joint anchor - com, child link frame */
  MBD_F_JQ = MBD_F_RC + 3,    /* float[4] joint frame rotation in the child link frame */
  MBD_F_RP = MBD_F_JQ + 4,    /* float[3] joint anchor - parent com, parent link frame */
  MBD_F_PQ = MBD_F_RP + 3,    /* float[4] joint frame rotation in the parent link frame */
  MBD_F_PARITY = MBD_F_PQ + 4,
  MBD_F_ANG_DAMP,      /* constraint_ang_damping */
  MBD_F_SLIDE,         /* int  bit k set: dof k of the link is a slide (prismatic) dof; only on links whose parent is the world */
  MBD_F_DOF0,          /* 3 x {stiffness, damping, lo, hi, act_id(int), gear, ctrl_lo, ctrl_hi} (lo/hi: radians, or metres for a slide) */
  MBD_F_NCON = MBD_F_DOF0 + MBD_MAXDOF * MBD_DOF_STRIDE, /* int */
  MBD_F_CON0,          /* MBD_MAXCON x {sx, sy, sz (sphere centre - com, link frame), radius, friction} */
  MBD_NFIELDS = MBD_F_CON0 + MBD_MAXCON * MBD_CON_STRIDE
};
enum { MBD_D_STIFF = 0, MBD_D_DAMP, MBD_D_LO, MBD_D_HI, MBD_D_ACT, MBD_D_GEAR, MBD_D_CLO, MBD_D_CHI };

#define MBD_BLOB_WORDS (MBD_HDR_WORDS + MBD_NFIELDS * MBD_MAXL)
#define MBD_LF(blob, f, l) ((blob)[MBD_HDR_WORDS + (f) * MBD_MAXL + (l)])

/* per-link recurrent state: x_i.pos(3) x_i.rot(4) xd_i.ang(3) xd_i.vel(3) */
#define MBD_STATE_STRIDE 13

#endif /* MBD_MODEL_H_ */
