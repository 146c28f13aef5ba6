/* Plain C99 caller of the C ABI (include/gr2m_b200.h): 3 subbasins x 24 months, one simulation and a
 * batch of 4 candidate parameter sets.  Build with `make demo`.  Without a CUDA device every compute
 * call reports GR2M_ERR_NODEVICE -- there is no CPU path -- and the program says so and exits 0. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "gr2m_b200.h"

int main(void)
{
    enum { NT = 24, NC = 3, NSETS = 4 };
    static double P[NC * NT], E[NC * NT], QS[NC * NT], qt[NT * NSETS], of[NSETS], qobs[NT];
    double area[NC] = {157.5, 248.5, 701.7};
    int ndays[NT], region[NC] = {0, 0, 0};
    double params[NSETS * 4] = {10.976, 0.665, 1.186, 1.169, 1000, 1, 1, 1, 300, 0.8, 1.1, 0.9, 50, 1.5, 0.9, 1.1};
    static const int dim[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    for (int t = 0; t < NT; ++t) {
        ndays[t] = dim[t % 12];
        for (int c = 0; c < NC; ++c) {
            P[c * NT + t] = floor(10.0 * (60.0 + 50.0 * sin(0.5 * t + c))) / 10.0;      /* column-major [NT x NC] */
            E[c * NT + t] = floor(10.0 * (120.0 + 30.0 * cos(0.5 * t))) / 10.0;
        }
    }
    printf("libgr2m_b200 ABI version %d, %d CUDA device(s)\n", gr2m_version(), gr2m_device_count());
    gr2m_basin *b = NULL;
    int rc = gr2m_basin_create(NT, NC, P, E, area, ndays, region, 1, &b);
    if (rc == GR2M_ERR_NODEVICE) {
        printf("no GPU here: %s\n", gr2m_last_error());
        return 0;
    }
    if (rc) { fprintf(stderr, "create failed (%d): %s\n", rc, gr2m_last_error()); return 1; }
    rc = gr2m_run(b, params, NULL, NULL, 0, NULL, NULL, NULL, NULL, QS, NULL, NULL);
    if (rc) { fprintf(stderr, "run failed (%d): %s\n", rc, gr2m_last_error()); return 1; }
    for (int t = 0; t < NT; ++t) qobs[t] = QS[t] + QS[NT + t] + QS[2 * NT + t];            /* set 0 is "truth" */
    rc = gr2m_objective_batch(b, params, NSETS, qobs, 6, GR2M_OF_NSE, 0, of, NULL, qt);
    if (rc) { fprintf(stderr, "objective failed (%d): %s\n", rc, gr2m_last_error()); return 1; }
    for (int s = 0; s < NSETS; ++s) printf("set %d: 1 - NSE = %.3f, outlet[0] = %.2f m3/s\n", s, of[s], qt[s * NT]);
    gr2m_basin_destroy(b);
    return of[0] == 0.0 ? 0 : 2;                                                           /* set 0 reproduces itself */
}
