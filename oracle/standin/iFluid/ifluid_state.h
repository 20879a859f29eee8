/* Stand-in for FronTier's ../iFluid/ifluid_state.h (collid.h:7) -- TEST INFRASTRUCTURE.
 * Fields are the ones documented at collid.h:20-39 plus the rigid-body fields the
 * sources use (dcollid.cpp:726-733,1083,1111; cdinit.cpp:192-204). */
#ifndef CLSN_ORACLE_IFLUID_STATE_STANDIN_H
#define CLSN_ORACLE_IFLUID_STATE_STANDIN_H

struct UF {
    POINT* next_pt;
    POINT* root;
    POINT* tail;
    int num_pts;
};

struct STATE {
    double vel[3];
    double collsnImpulse[3];
    double collsnImpulse_RG[3];
    double friction[3];
    double avgVel[3];
    double x_old[3];
    int collsn_num;
    int collsn_num_RG;
    bool has_collsn;
    bool is_fixed;
    bool is_movableRG;
    UF impZone;
};

#endif
