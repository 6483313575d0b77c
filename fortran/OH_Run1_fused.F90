! OH_Run1_fused.F90 — the body a patched OH_GridCompMod Run1 uses instead of
! PREP_FOR_BOOST / CALL_BOOST / the troposphere WHERE / the unit conversion
! (reference: OH_GridComp/OH_GridCompMod.F90:1232-1599).  Source text only: this build box has no
! Fortran compiler and no ESMF/MAPL (SURVEY.md section 0); the executable equivalent is
! quickchem_b200/capi.py::OhRun1 driving the same C entry points, exercised by tests/test_gpu_run1.py.
!
! What stays in Fortran: the MAPL surface (SetServices / Initialize / Run / Run2, the alarm gate
! :1180-1185, need_to_call_BOOST :1189-1193), import selection per OH_data_source (:1310-1436,
! :1493-1540) — it only decides WHICH import pointer is handed over — and the DIAG_* exports.
! What moves to the GPU behind one call: PL/TV/NDWET (:1247-1257), latarr, stratO3,
! gridBoxThickness, aod (:1444-1466), the six vertical sums (:1468-1478), noon SZA (:1481-1482),
! the 27-feature pack (:303-345), XGDMatrixCreateFromMat + XGBoosterPredict (:347-356),
! 10**x (:369), *OHscale (:1569), the WHERE mask (:1579-1587) and *NDWET*1e-6 (:1595).
!
!   use xgb_fortran_api          ! unchanged reference module, now resolved by libqcoh.so
!   use qcoh_fortran_api
!
!   type(c_ptr),          save :: xx_bst               ! as the reference's SAVE booster (:182)
!   type(c_ptr),          save :: oh_dev = c_null_ptr  ! fused handle, one per process / GPU
!   logical,              save :: first_time = .TRUE.
!   type(qcoh_oh_config)       :: cfg
!   type(qcoh_run1_in)         :: rin
!   type(qcoh_run1_out)        :: rout
!   type(c_ptr)                :: dummy
!
!   ONE_TIME_SETUP: IF ( first_time ) THEN
!      rc = XGBoosterCreate_f( dummy, 0_c_int64_t, xx_bst )              ! (:256) len = 0: dmats unused
!      _ASSERT(rc==0,'Failed in XGBoosterCreate_f')
!      rc = XGBoosterLoadModel_f( xx_bst, XGBoostFilename )              ! (:261)
!      _ASSERT(rc==0,'Failed in XGBoosterLoadModel_f: '//qcoh_last_error())
!      cfg%ncol = im*jm ;  cfg%km = km
!      cfg%mapl_epsilon = MAPL_EPSILON ;  cfg%mapl_avogad = MAPL_AVOGAD ;  cfg%mapl_runiv = MAPL_RUNIV
!      cfg%mapl_radians_to_degrees = MAPL_RADIANS_TO_DEGREES
!      cfg%mapl_degrees_to_radians = MAPL_DEGREES_TO_RADIANS
!      cfg%ohscale = self%OHscale
!      cfg%compute_once_per_day = merge(1, 0, self%compute_once_per_day)
!      cfg%tropp_min = 40.0 * 100                                        ! (:1563)
!      cfg%missing = -999.0                                              ! (:213)
!      rc = qcoh_oh_create( xx_bst, cfg, oh_dev )
!      _ASSERT(rc==0,'Failed in qcoh_oh_create: '//qcoh_last_error())
!      first_time = .FALSE.
!   END IF ONE_TIME_SETUP
!
!   ! optional, off by default (new rc key `reload_model_on_month_change: F`): follow the %m2 of the file
!   ! pattern instead of keeping the start month's booster for the whole run (reference behaviour, :182,209)
!   IF ( need_to_call_BOOST .AND. self%reload_model_on_month_change ) THEN
!      rc = qcoh_oh_select_model( oh_dev, TRIM(self%XGBoostFilePattern)//c_null_char, nymd, nhms, changed )
!      _ASSERT(rc==0, qcoh_last_error())
!   END IF
!
!   rin%nymd = nymd
!   rin%need_to_call_boost = merge(1, 0, need_to_call_BOOST)
!   rin%T_MOD   = c_loc(T_MOD)   ;  rin%Q_MOD  = c_loc(Q_MOD)
!   rin%PLE_MOD = c_loc(PLE_MOD) ;  rin%TROPP  = c_loc(TROPP_MOD)
!   rin%T_BST   = c_loc(bb%T)    ;  rin%Q_BST  = c_loc(bb%QV)     ! whichever import :1326-1343 picked
!   rin%PLE_BST = c_loc(PLE_BST) ;  rin%ZLE_BST = c_loc(ZLE_BST)
!   rin%TAUCLW  = c_loc(TAUCLW)  ;  rin%TAUCLI = c_loc(TAUCLI)
!   rin%FCLD    = c_loc(bb%CLOUD);  rin%CH4 = c_loc(bb%CH4) ;  rin%CO = c_loc(bb%CO)
!   rin%SCA(1)  = c_loc(BCscacoef_4D(:,:,:,self%wavelength_index))   ! contiguous 3-D slab; or the _3D array
!   ...          (OC, BR, DU, SU, SS, NI likewise, order of :1456-1465)
!   rin%NO2 = c_loc(oh_NO2) ; rin%O3 = c_loc(oh_O3) ; ... ; rin%CH2O = c_loc(oh_CH2O)
!   rin%GMITO3 = c_loc(oh_GMITO3) ; rin%GMITTO3 = c_loc(oh_GMITTO3) ; rin%ALBUV = c_loc(oh_ALBUV)
!   rin%LATS = c_loc(LATS) ; rin%LONS = c_loc(LONS) ; rin%OH_CLIM = c_loc(default_OH)
!   rin%AREA = c_null_ptr
!
!   rout%OH = c_loc(OH) ;  rout%OH_boost = c_null_ptr ;  rout%NDWET = c_null_ptr
!   CALL MAPL_GetPointer(export, ptr3d, 'OH_boost', __RC__)
!   IF (ASSOCIATED(ptr3d)) rout%OH_boost = c_loc(ptr3d)              ! (:1571-1572)
!   CALL MAPL_GetPointer(export, ptr3d, 'DIAG_NDWET', __RC__)
!   IF (ASSOCIATED(ptr3d)) rout%NDWET = c_loc(ptr3d)                 ! (:1598-1599)
!   rout%X = c_null_ptr ;  rout%pred = c_null_ptr
!   rout%LOSS_CH4 = c_null_ptr ;  rout%LOSS_CO = c_null_ptr          ! or c_loc of a new export, for CH4 / CO
!
!   rc = qcoh_oh_run1( oh_dev, rin, rout )
!   _ASSERT(rc==0, qcoh_last_error())      ! carries 'Minimum tropopause pressure is not low enough!' (:288)
!
!   AFTER_BOOST: IF ( need_to_call_BOOST ) THEN      ! diagnostics (:1602-1728) straight from HBM
!      CALL MAPL_GetPointer(export, ptr3d, 'DIAG_AODUP', __RC__)
!      IF (ASSOCIATED(ptr3d)) rc = qcoh_oh_get_diag( oh_dev, 'AODUP'//c_null_char, c_loc(ptr3d) )
!      ...  TAUCLWDN TAUCLIDN TAUCLIUP TAUCLWUP AODDN PL NDWET OH_boost (3-D), LAT SZA stratO3 (2-D);
!      the remaining DIAG_* exports are plain copies of import fields and stay in Fortran
!   END IF AFTER_BOOST
!
! The persistent self%OH_ML(:,:,:) (:76-78,:893) lives in HBM inside oh_dev; with
! compute_once_per_day the 23 non-boost steps of a day upload only T, Q, PLE, TROPP and oh_OH.
