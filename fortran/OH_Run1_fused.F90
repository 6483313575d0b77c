! OH_Run1_fused.F90 — the fused, device-resident replacement of the body of OH_GridCompMod::Run1,
! i.e. of OH_GridComp/OH_GridCompMod.F90:1226-1740: model-state PL/TV/NDWET, PREP_FOR_BOOST (import selection,
! derived 2-D / 3-D fields, vertical sums, noon SZA), CALL_BOOST (predict_OH_with_XGB, * OHscale, OH_boost export),
! the troposphere WHERE, the unit conversion, and the DIAG_* exports.
!
! How it is used (fortran/OH_GridCompMod.F90.patch): this file is #include'd in the CONTAINS section of
! OH_GridCompMod, so that it sees type(OH_GridComp), the OH_data_source constants and MAPL / ESMF exactly as Run1
! does; Run1 calls Run1_fused right after it has the dimensions (:1203) when built with -DQCOH_FUSED_RUN1.
! Everything numerical happens in libqcoh.so on the GPU behind qcoh_oh_run1 (include/qcoh.h); what stays in
! Fortran is what only MAPL can do: fetching the import / export pointers by name.
!
! Not compiled on the build box (no Fortran compiler, no ESMF / MAPL there): tests/test_gpu_model_day.py drives
! the same C entry points in the same order through ctypes (24 hourly steps, all three OH_data_source modes).

  subroutine Run1_fused (self, import, export, XGBoostFilename, nymd, nhms, im, jm, km, &
                         need_to_call_BOOST, LATS, LONS, OH, rc)
    use iso_c_binding
    use xgb_fortran_api      ! the reference's own interface module, unchanged: resolved by libqcoh.so
    use qcoh_fortran_api     ! fortran/qcoh_fortran_api.F90

    type (OH_GridComp), intent(inout)          :: self
    type (ESMF_State),  intent(inout)          :: import, export
    character (len=*),  intent(in)             :: XGBoostFilename
    integer,            intent(in)             :: nymd, nhms, im, jm, km
    logical,            intent(in)             :: need_to_call_BOOST
    real, pointer,      intent(in)             :: LATS(:,:), LONS(:,:)
    real, pointer                              :: OH(:,:,:)       ! INTERNAL state, molec/cm3 (OH_StateSpecs.rc:84)
    integer, optional,  intent(out)            :: rc

    ! one booster and one fused handle per process, like the reference's SAVE'd xx_bst (:182, :209)
    type (c_ptr), save          :: xx_bst = c_null_ptr
    type (c_ptr), save          :: oh_dev = c_null_ptr
    logical,      save          :: first_time = .TRUE.

    type (qcoh_oh_config), target :: cfg
    type (qcoh_run1_in),   target :: rin
    type (qcoh_run1_out),  target :: rout
    type (c_ptr)                  :: no_dmats
    integer (c_int)               :: crc, changed

    real, pointer, dimension(:,:,:)   :: T_MOD, Q_MOD, PLE_MOD
    real, pointer, dimension(:,:)     :: TROPP_MOD
    real, pointer, dimension(:,:,:)   :: T_BST, Q_BST, PLE_BST, ZLE_BST, TAUCLW, TAUCLI, FCLD, CH4, CO
    real, pointer, dimension(:,:,:)   :: sca3
    real, pointer, dimension(:,:,:,:) :: sca4
    real, pointer, dimension(:,:,:)   :: gas, default_OH, ptr3d, t_avg
    real, pointer, dimension(:,:)     :: GMITO3, GMITTO3, ALBUV, ptr2d
    real, pointer, dimension(:,:,:)   :: oh_boost_exp, ndwet_exp

    character (len=2), parameter :: species(7) = (/ 'BC', 'OC', 'BR', 'DU', 'SU', 'SS', 'NI' /)   ! order of :1456-1465
    character (len=4), parameter :: gases(11)  = (/ 'NO2 ', 'O3  ', 'ISOP', 'ACET', 'C2H6', 'C3H8', &
                                                    'PRPE', 'ALK4', 'MP  ', 'H2O2', 'CH2O' /)
    type (c_ptr) :: gas_ptr(11)
    integer :: s, g, STATUS
    character (len=ESMF_MAXSTR) :: Iam

    Iam = 'OH::Run1_fused'

    ! ---- one-time set-up: booster (:242-271 without the dummy DMatrix: nothing needs it) + the fused handle
    if (first_time) then
       no_dmats = c_null_ptr
       crc = XGBoosterCreate_f (no_dmats, 0_c_int64_t, xx_bst)                  ! len = 0: dmats is never read
       _ASSERT(crc == 0, 'Failed in XGBoosterCreate_f')
       crc = XGBoosterLoadModel_f (xx_bst, XGBoostFilename)
       _ASSERT(crc == 0, 'Failed in XGBoosterLoadModel_f: '//qcoh_last_error())
       cfg%ncol = im*jm
       cfg%km   = km
       cfg%mapl_epsilon = MAPL_EPSILON
       cfg%mapl_avogad  = MAPL_AVOGAD
       cfg%mapl_runiv   = MAPL_RUNIV
       cfg%mapl_radians_to_degrees = MAPL_RADIANS_TO_DEGREES
       cfg%mapl_degrees_to_radians = MAPL_DEGREES_TO_RADIANS
       cfg%ohscale = self%OHscale
       cfg%compute_once_per_day = merge (1, 0, self%compute_once_per_day)    ! dynamic_k_range = .NOT. this (:1561)
       cfg%tropp_min = 40.0 * 100                                            ! hPa -> Pa (:1563)
       cfg%missing   = -999.0                                                ! xx_miss (:213)
       crc = qcoh_oh_create (xx_bst, cfg, oh_dev)
       _ASSERT(crc == 0, 'Failed in qcoh_oh_create: '//qcoh_last_error())
       first_time = .FALSE.
    end if

    ! optional, off unless the rc file says `reload_model_on_month_change: .TRUE.`: follow the %m2 of the file
    ! pattern instead of predicting with the start month's booster for the whole run (reference behaviour)
    if (need_to_call_BOOST .and. self%reload_model_on_month_change) then
       crc = qcoh_oh_select_model (oh_dev, trim(self%XGBoostFilePattern)//c_null_char, nymd, nhms, changed)
       _ASSERT(crc == 0, 'Failed in qcoh_oh_select_model: '//qcoh_last_error())
    end if

    ! ---- current model state (:1233-1236); PL_MOD, TV_MOD, NDWET_MOD (:1247-1257) are computed on the device
    call MAPL_GetPointer (import, T_MOD,     'T',     __RC__)
    call MAPL_GetPointer (import, Q_MOD,     'Q',     __RC__)
    call MAPL_GetPointer (import, PLE_MOD,   'PLE',   __RC__)
    call MAPL_GetPointer (import, TROPP_MOD, 'TROPP', __RC__)
    _ASSERT(lbound(PLE_MOD,3) == 0, 'Error. Expecting PLE starting index 0')

    rin%nymd = nymd
    rin%need_to_call_boost = merge (1, 0, need_to_call_BOOST)
    rin%T_MOD   = first3 (T_MOD)
    rin%Q_MOD   = first3 (Q_MOD)
    rin%PLE_MOD = first3 (PLE_MOD)
    rin%TROPP   = first2 (TROPP_MOD)
    rin%AREA    = c_null_ptr           ! the global-mean diagnostic is not part of the reference

    ! default OH above the tropopause (:1548), also the DIAG_OH_M2G export (:1553-1554)
    call MAPL_GetPointer (import, default_OH, 'oh_OH', __RC__)
    rin%OH_CLIM = first3 (default_OH)
    call MAPL_MaxMin ('OH: OH From M2G ', default_OH)
    call MAPL_GetPointer (export, ptr3d, 'DIAG_OH_M2G', __RC__)
    if (associated(ptr3d)) ptr3d(:,:,:) = default_OH(:,:,:)

    ! ZLE and CH4 are also read by non-boost steps when the diagnostic is on; hand them over on boost steps only
    rin%ZLE_BST = c_null_ptr
    rin%CH4     = c_null_ptr

    if (need_to_call_BOOST) then
       ! the 24-hour-average spin-up switch (:1307-1317)
       self%use_inst_values = .FALSE.
       if (self%OH_data_source == ONLINE_AVG24) then
          call MAPL_GetPointer (import, t_avg, 'T_avg24', __RC__)
          if (t_avg(1,1,1) == 0.0) self%use_inst_values = .TRUE.
       end if
       if (mapl_am_i_root()) then
          if (       self%use_inst_values) print *, 'OH is in the SPINUP period for 24-hour averages'
          if (.not.  self%use_inst_values) print *, 'OH is *NOT* in the SPINUP period for 24-hour averages'
       end if

       ! the values handed to boost, per OH_data_source (:1326-1372, :1493-1525)
       call MAPL_GetPointer (import, T_BST,   trim(boost_import(self, 'T')),      __RC__)
       call MAPL_GetPointer (import, Q_BST,   trim(boost_import(self, 'Q')),      __RC__)
       call MAPL_GetPointer (import, PLE_BST, trim(boost_import(self, 'PLE')),    __RC__)
       call MAPL_GetPointer (import, ZLE_BST, trim(boost_import(self, 'ZLE')),    __RC__)
       call MAPL_GetPointer (import, TAUCLW,  trim(boost_import(self, 'TAUCLW')), __RC__)
       call MAPL_GetPointer (import, TAUCLI,  trim(boost_import(self, 'TAUCLI')), __RC__)
       call MAPL_GetPointer (import, FCLD,    trim(boost_import(self, 'FCLD')),   __RC__)
       call MAPL_GetPointer (import, CH4,     trim(boost_import(self, 'CH4')),    __RC__)
       call MAPL_GetPointer (import, CO,      trim(boost_import(self, 'CO')),     __RC__)
       _ASSERT(lbound(PLE_BST,3) == 0, 'Error. Expecting PLE starting index 0')
       _ASSERT(lbound(ZLE_BST,3) == 0, 'Error. Expecting ZLE starting index 0')
       rin%T_BST   = first3 (T_BST)
       rin%Q_BST   = first3 (Q_BST)
       rin%PLE_BST = first3 (PLE_BST)
       rin%ZLE_BST = first3 (ZLE_BST)
       rin%TAUCLW  = first3 (TAUCLW)
       rin%TAUCLI  = first3 (TAUCLI)
       rin%FCLD    = first3 (FCLD)
       rin%CH4     = first3 (CH4)
       rin%CO      = first3 (CO)

       ! scattering coefficients: archived fields are 3-D, online ones 4-D with a wavelength axis (:1388-1436);
       ! the slab (:,:,:,wavelength_index) is contiguous, its first element is handed over
       do s = 1, 7
          if (self%OH_data_source == PRECOMPUTED) then
             call MAPL_GetPointer (import, sca3, 'oh_'//species(s)//'SCACOEF', __RC__)
             rin%SCA(s) = first3 (sca3)
          else
             call MAPL_GetPointer (import, sca4, trim(boost_import(self, species(s)//'SCACOEF')), __RC__)
             rin%SCA(s) = c_loc (sca4(lbound(sca4,1), lbound(sca4,2), lbound(sca4,3), self%wavelength_index))
          end if
       end do

       ! always-climatological gases (:1491-1492, :1508-1515, :1539) in the order of qcoh_run1_in
       do g = 1, 11
          call MAPL_GetPointer (import, gas, 'oh_'//trim(gases(g)), __RC__)
          gas_ptr(g) = first3 (gas)
       end do
       rin%NO2  = gas_ptr(1);  rin%O3   = gas_ptr(2);  rin%ISOP = gas_ptr(3);  rin%ACET = gas_ptr(4)
       rin%C2H6 = gas_ptr(5);  rin%C3H8 = gas_ptr(6);  rin%PRPE = gas_ptr(7);  rin%ALK4 = gas_ptr(8)
       rin%MP   = gas_ptr(9);  rin%H2O2 = gas_ptr(10); rin%CH2O = gas_ptr(11)

       call MAPL_GetPointer (import, GMITO3,  'oh_GMITO3',  __RC__)
       call MAPL_GetPointer (import, GMITTO3, 'oh_GMITTO3', __RC__)
       call MAPL_GetPointer (import, ALBUV,   'oh_ALBUV',   __RC__)
       rin%GMITO3  = first2 (GMITO3)
       rin%GMITTO3 = first2 (GMITTO3)
       rin%ALBUV   = first2 (ALBUV)
       rin%LATS    = first2 (LATS)
       rin%LONS    = first2 (LONS)
    end if

    ! ---- outputs: the INTERNAL OH always; OH_boost on boost steps (:1571-1572); DIAG_NDWET every step (:1598-1599)
    rout%OH       = first3 (OH)
    rout%OH_boost = c_null_ptr
    rout%NDWET    = c_null_ptr
    rout%X        = c_null_ptr
    rout%pred     = c_null_ptr
    rout%LOSS_CH4 = c_null_ptr
    rout%LOSS_CO  = c_null_ptr
    if (need_to_call_BOOST) then
       call MAPL_GetPointer (export, oh_boost_exp, 'OH_boost', __RC__)
       if (associated(oh_boost_exp)) rout%OH_boost = first3 (oh_boost_exp)
    end if
    call MAPL_GetPointer (export, ndwet_exp, 'DIAG_NDWET', __RC__)
    if (associated(ndwet_exp)) rout%NDWET = first3 (ndwet_exp)

    ! ---- the hot path: assembly -> predict -> 10**x * OHscale -> WHERE(PL > TROPP) -> * NDWET * 1e-6
    crc = qcoh_oh_run1 (oh_dev, rin, rout)
    _ASSERT(crc == 0, 'OH Prediction: '//qcoh_last_error())

    ! ---- diagnostics that are only meaningful on a boost step (:1602-1735)
    if (need_to_call_BOOST) then
       ! derived on the device: copied out of HBM on demand
       call diag2 ('DIAG_LAT',        'LAT')
       call diag3 ('DIAG_TAUCLWDN',   'TAUCLWDN')
       call diag3 ('DIAG_TAUCLIDN',   'TAUCLIDN')
       call diag3 ('DIAG_TAUCLIUP',   'TAUCLIUP')
       call diag3 ('DIAG_TAUCLWUP',   'TAUCLWUP')
       call diag2 ('DIAG_GMISTRATO3', 'stratO3')
       call diag3 ('DIAG_AODUP',      'AODUP')
       call diag3 ('DIAG_AODDN',      'AODDN')
       call diag2 ('DIAG_SZA',        'SZA')
       call diag3 ('DIAG_PL',         'PL')
       call diag3 ('DIAG_AOD',        'AOD')
       ! plain copies of what was handed to boost
       call MAPL_GetPointer (export, ptr2d, 'DIAG_ALBUV', __RC__)
       if (associated(ptr2d)) ptr2d(:,:)   = ALBUV
       call MAPL_GetPointer (export, ptr3d, 'DIAG_T',     __RC__)
       if (associated(ptr3d)) ptr3d(:,:,:) = T_BST
       call MAPL_GetPointer (export, ptr3d, 'DIAG_CH4',   __RC__)
       if (associated(ptr3d)) ptr3d(:,:,:) = CH4
       call MAPL_GetPointer (export, ptr3d, 'DIAG_CO',    __RC__)
       if (associated(ptr3d)) ptr3d(:,:,:) = CO
       call MAPL_GetPointer (export, ptr3d, 'DIAG_CLOUD', __RC__)
       if (associated(ptr3d)) ptr3d(:,:,:) = FCLD
       call MAPL_GetPointer (export, ptr3d, 'DIAG_QV',    __RC__)
       if (associated(ptr3d)) ptr3d(:,:,:) = Q_BST
       call MAPL_GetPointer (export, ptr3d, 'DIAG_ZLE',   __RC__)
       if (associated(ptr3d)) ptr3d(:,:,:) = ZLE_BST
       call MAPL_GetPointer (export, ptr3d, 'DIAG_C2H6',  __RC__)
       if (associated(ptr3d)) then
          call MAPL_GetPointer (import, gas, 'oh_C2H6', __RC__)
          ptr3d(:,:,:) = gas
       end if
       call MAPL_GetPointer (export, ptr3d, 'DIAG_ISOP',  __RC__)
       if (associated(ptr3d)) then
          call MAPL_GetPointer (import, gas, 'oh_ISOP', __RC__)
          ptr3d(:,:,:) = gas
       end if
       do s = 1, 7
          call MAPL_GetPointer (export, ptr3d, 'DIAG_SC_'//species(s), __RC__)
          if (associated(ptr3d)) then
             if (self%OH_data_source == PRECOMPUTED) then
                call MAPL_GetPointer (import, sca3, 'oh_'//species(s)//'SCACOEF', __RC__)
                ptr3d(:,:,:) = sca3
             else
                call MAPL_GetPointer (import, sca4, trim(boost_import(self, species(s)//'SCACOEF')), __RC__)
                ptr3d(:,:,:) = sca4(:,:,:,self%wavelength_index)
             end if
          end if
       end do
    end if

    RETURN_(ESMF_SUCCESS)

  contains

    ! address of the first element of a MAPL pointer array, whatever its lower bounds (PLE / ZLE start at 0)
    function first3 (a) result (p)
      real, pointer, intent(in) :: a(:,:,:)
      type (c_ptr) :: p
      p = c_loc (a(lbound(a,1), lbound(a,2), lbound(a,3)))
    end function first3

    function first2 (a) result (p)
      real, pointer, intent(in) :: a(:,:)
      type (c_ptr) :: p
      p = c_loc (a(lbound(a,1), lbound(a,2)))
    end function first2

    ! DIAG_<export> <- the derived field `field` of the last boost step, if the export is wanted
    subroutine diag3 (export_name, field)
      character (len=*), intent(in) :: export_name, field
      real, pointer :: e(:,:,:)
      integer :: STATUS
      call MAPL_GetPointer (export, e, export_name, RC=STATUS)
      if (STATUS /= 0 .or. .not. associated(e)) return
      STATUS = qcoh_oh_get_diag (oh_dev, field//c_null_char, first3 (e))
    end subroutine diag3

    subroutine diag2 (export_name, field)
      character (len=*), intent(in) :: export_name, field
      real, pointer :: e(:,:)
      integer :: STATUS
      call MAPL_GetPointer (export, e, export_name, RC=STATUS)
      if (STATUS /= 0 .or. .not. associated(e)) return
      STATUS = qcoh_oh_get_diag (oh_dev, field//c_null_char, first2 (e))
    end subroutine diag2

  end subroutine Run1_fused


  ! Which import feeds a boost-state field (OH_GridCompMod.F90:1326-1436, :1493-1525): the archived 'oh_X' when
  ! PRECOMPUTED, the instantaneous 'X' when ONLINE_INST or during the 24-hour-average spin-up, else 'X_avg24'.
  ! (libqcoh restates the same table for non-Fortran hosts: qcoh_import_name, run1_control.cpp.)
  function boost_import (self, base) result (name)
    type (OH_GridComp), intent(in) :: self
    character (len=*),  intent(in) :: base
    character (len=ESMF_MAXSTR)    :: name
    select case (self%OH_data_source)
    case (PRECOMPUTED)
       name = 'oh_'//base
    case (ONLINE_AVG24)
       if (self%use_inst_values) then
          name = base
       else
          name = trim(base)//'_avg24'
       end if
    case default        ! ONLINE_INST
       name = base
    end select
  end function boost_import
