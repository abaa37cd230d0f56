/*
 * rt2025_rng.h — the sampling contract shared by the CUDA core and the CPU oracle.
 *
 * The reference draws every random number from the unseeded thread-local `rand::rng()`
 * (src/utils/random.rs:8-14), so its sample stream is neither reproducible nor ordered.
 * Both implementations here replace it with the counter-based Philox4x32-10 generator
 * (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11):
 *
 *     key     = (seed_lo, seed_hi)
 *     counter = (pixel_index, sample_index, segment, slot)
 *
 * pixel_index = j*image_width + i, sample_index = s_i*sqrt_spp + s_j (camera.rs:183-184),
 * segment = number of world.hit calls made before this one on the path (0 = camera ray),
 * slot = which decision of that segment consumes the numbers (below).  One Philox call yields
 * four 32-bit words x0..x3 and two doubles in [0,1):
 *
 *     a = ((x0 >> 5) * 2^26 + (x1 >> 6)) * 2^-53,   b = ((x2 >> 5) * 2^26 + (x3 >> 6)) * 2^-53
 *
 * Because a draw is addressed, not sequenced, results do not depend on traversal order,
 * wavefront scheduling, GPU count or on how many draws another branch consumed.
 */
#ifndef RT2025_RNG_H
#define RT2025_RNG_H

#define RT_PHILOX_M0 0xD2511F53u
#define RT_PHILOX_M1 0xCD9E8D57u
#define RT_PHILOX_W0 0x9E3779B9u
#define RT_PHILOX_W1 0xBB67AE85u

/* slots — the reference call that consumes each one */
#define RT_SLOT_CAM_JITTER 0u  /* a,b = sample_square_stratified draws       camera.rs:263-268 */
#define RT_SLOT_CAM_DISK 1u    /* a = theta/(2pi), b = r^2  random_in_unit_disk vec3.rs:63-69   */
#define RT_SLOT_CAM_TIME 2u    /* a = ray_time                               camera.rs:258     */
#define RT_SLOT_MATERIAL 3u    /* Metal: a,b = random_unit_vector r1,r2 (material.rs:85);      */
                               /* Dielectric: a = reflectance test (material.rs:129)            */
#define RT_SLOT_MIXTURE 4u     /* a = MixturePDF pick (pdf.rs:114); b = light leaf pick          */
                               /*     (Hittables::random, hits.rs:69-75)                        */
#define RT_SLOT_DIRECTION 5u   /* a,b = r1,r2 of the chosen generator: random_cosine_direction  */
                               /* (vec3.rs:333), random_unit_vector (vec3.rs:313), Quad::random  */
                               /* (quad.rs:122), random_to_sphere (sphere.rs:63), Triangle::random */
#define RT_SLOT_MIX 6u         /* a = Mix material pick (material.rs:251); nesting level L uses  */
                               /*     slot 6 + 256*L                                            */
#define RT_SLOT_DISNEY 7u      /* a = lobe pick (disney.rs:674), b = the lobe's own extra decision:     */
                               /*     diff_trans flip (disney.rs:607) or Fresnel pick (disney.rs:651);   */
                               /*     the lobe's r0,r1 are RT_SLOT_DIRECTION                             */
#define RT_SLOT_MEDIUM0 16u    /* free-flight draw of medium m (volume.rs:58), m = index into rt_scene_desc.media:          */
                               /*     slot 16 + m/2, component a for even m, b for odd m - one Philox call serves two media */
#define RT_MEDIUM_SLOT(m) (RT_SLOT_MEDIUM0 + ((m) >> 1))

/* light leaf pick: the lights tree is flattened depth first; leaf i carries the probability
 * w_i = product over its ancestors of 1/len (hits.rs:69-75 picks uniformly at every level) and
 * cdf_i = w_0 + ... + w_i accumulated in that order in binary64.  A draw x selects the first
 * leaf with x < cdf_i (the last leaf if none). */

#endif
